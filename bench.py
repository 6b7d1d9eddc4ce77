#!/usr/bin/env python
"""Headline benchmark of the hot path (BASELINE.json): volumes/sec of nnU-Net 3d_fullres sliding-window
inference with 8x mirror TTA + Gaussian aggregation on a synthetic 182x218x182 FLAIR volume,
Generic_UNet random-init, patch 128^3 (config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1: cohort sharding, every rank runs K subjects, no collective on
the data path -> "scaling": "weak").  A step = one volume through z-score -> 12 tiles x 8 mirrors of
Generic_UNet -> overlap-add -> finalize.  `value` is timed with the raw volume already in HBM; `e2e` goes
through the C-ABI host-buffer call (dwmh_predict_volume_host) with pinned host memory, H2D + D2H inside.
`--impl reference` times the CPU oracle (the reference's PyTorch arithmetic, oracle/) on the host cores.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (182, 218, 182)
N_TILES, N_MIRRORS = 12, 8
WORKLOAD = "synthetic 182x218x182 FLAIR, 3d_fullres Generic_UNet random-init (31.2M params), patch 128^3, step 0.5, 8x mirror TTA, Gaussian aggregation, z-score(nonzero mask)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if mx and s > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_forward_sample(n_forwards, threads):
    """Time n tile-forwards (+softmax) of the TTA workload with the oracle; returns seconds per forward."""
    import torch
    import oracle as O
    torch.set_num_threads(threads)
    net = O.build_benchmark_network(0)
    raw = O.synthetic_flair(SHAPE, seed=0)
    data = raw.copy()
    data[0] = O.zscore_nnunet(raw[0], np.where(raw[0] != 0, 0, -1), True)
    tile = torch.from_numpy(np.ascontiguousarray(data[None, :, 0:128, 0:128, 0:128]))
    g = torch.from_numpy(O.get_gaussian((128,) * 3))
    times = []
    for m in range(n_forwards):
        dims = O.MIRROR_DIMS[m % 8]
        t0 = time.perf_counter()
        with torch.no_grad():
            x = torch.flip(tile, dims) if dims else tile
            p = torch.softmax(net(x), 1)
            p = torch.flip(p, dims) if dims else p
            _ = p * (1.0 / 8) * g
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    times = cpu_forward_sample(args.warmup + args.steps, cores)[args.warmup:]
    per_fwd = sum(times) / len(times)
    vps = 1.0 / (per_fwd * N_TILES * N_MIRRORS)
    sample = "each step = 1 of the 96 tile-forwards (128^3 patch, mirror m = step mod 8, softmax, x Gaussian/8) of the workload, fp32 oracle, torch CPU %d threads; volumes/s = 1/(96 x s per forward)" % cores
    print(json.dumps({
        "impl": "reference", "metric": "volumes/sec", "value": vps, "unit": "volumes/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_fwd * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "step_definition": "1/96 volume (one tile-forward)"},
        "cpu_baseline": {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": vps, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def conv_algorithmic_bytes(plans, raw32_max_edge=64):
    """Bytes every conv / transposed-conv launch of ONE tile-forward must move at least once: input activations
    (fp16; fp32 for the 1-channel input volume), output (fp16 raw, fp32 for layers whose output edge <= raw32_max_edge)
    and the layer's fp16 weights.  Returns (total bytes, number of launches)."""
    import numpy as np
    st = plans["plans_per_stage"][max(plans["plans_per_stage"].keys())]
    pools, kers = st["pool_op_kernel_sizes"], st["conv_kernel_sizes"]
    shape = np.array(st["patch_size"], dtype=np.int64)
    feats = [plans["base_num_features"]]
    for _ in pools:
        feats.append(min(int(round(feats[-1] * 2)), 320))
    total, launches, cin, shapes = 0, 0, plans["num_modalities"], []
    in_shape = shape.copy()

    def out_bytes(c, sp, last=False):
        return int(np.prod(sp)) * c * (4 if (max(sp) <= raw32_max_edge and not last) else 2)
    for d in range(len(pools) + 1):
        if d > 0:
            shape = shape // np.array(pools[d - 1])
        k = int(np.prod(kers[d]))
        total += int(np.prod(in_shape)) * cin * (4 if d == 0 else 2) + out_bytes(feats[d], shape) + cin * feats[d] * k * 2
        total += int(np.prod(shape)) * feats[d] * 2 + out_bytes(feats[d], shape) + feats[d] * feats[d] * k * 2
        launches += 2
        cin, in_shape = feats[d], shape.copy()
        shapes.append(shape.copy())
    for u in range(len(pools)):
        skip, sp = feats[-(2 + u)], shapes[-(2 + u)]
        k = int(np.prod(kers[-(u + 1)]))
        total += int(np.prod(in_shape)) * cin * 2 + int(np.prod(sp)) * skip * 2 + cin * skip * int(np.prod(pools[-(u + 1)])) * 2      # transposed conv
        total += int(np.prod(sp)) * 2 * skip * 2 + out_bytes(skip, sp) + 2 * skip * skip * k * 2
        total += int(np.prod(sp)) * skip * 2 + out_bytes(skip, sp, last=(u == len(pools) - 1)) + skip * skip * k * 2
        launches += 3
        cin, in_shape = skip, sp.copy()
    return total, launches


def run_b200(args):
    import torch
    import torch.distributed as dist
    import deepwmh_b200
    from deepwmh_b200 import workload as W      # synthetic volume + random-init weights (product side; oracle/ is only the cpu_baseline leg)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pk = peaks()

    plans = deepwmh_b200.benchmark_plans()
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=local, act_dtype=args.dtype, max_batch=args.max_batch)
    tr.load_checkpoint_ram({"state_dict": W.random_init_state_dict(plans, 0)}, False)
    n = tr.network
    flops_fwd = W.forward_flops(plans)

    # per-rank subject (cohort sharding: rank r owns subjects r, r+world, ...)
    raws = [W.synthetic_flair(SHAPE, seed=rank + world * i)[0] for i in range(2)]
    raw_dev = [torch.from_numpy(r).to(dev) for r in raws]
    pinned = [torch.from_numpy(r).pin_memory() for r in raws]
    X, Y, Z = SHAPE
    V = X * Y * Z
    vol = torch.empty(SHAPE, dtype=torch.float32, device=dev)
    agg = torch.empty((2,) + SHAPE, dtype=torch.float32, device=dev)
    wgt = torch.empty(SHAPE, dtype=torch.float32, device=dev)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    sm_out = torch.empty((2,) + SHAPE, dtype=torch.float32).pin_memory()
    seg_out = torch.empty(SHAPE, dtype=torch.uint8).pin_memory()

    def step_device(i):
        flush.zero_()
        vol.copy_(raw_dev[i % 2])                  # device-to-device: input resident in HBM
        n.normalize_(vol, None, 2)
        agg.zero_(); wgt.zero_()
        n.accumulate_tiles(vol, agg, wgt, 0.5, True, (0, 1, 2), True)
        return n.finalize(agg, wgt)

    def step_e2e(i):
        flush.zero_()
        tr.predict_raw_volume_host(pinned[i % 2].numpy(), 2, True, (0, 1, 2), 0.5, True, sm_out.numpy(), seg_out.numpy())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        k0, _ = n.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        k1, _ = n.counters()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, k1 - k0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, max(1, args.steps), 1)

    # roofline of the dominant kernel family (conv stack): device time of the conv stack, CUDA events
    # recorded by the library on the launching stream around the 96 forwards of one more volume
    n._lib.dwmh_set_stage_timing(n._ctx, 1)
    step_device(0); torch.cuda.synchronize()
    import ctypes as C
    st = (C.c_float * 4)()
    n._lib.dwmh_get_stage_timing(n._ctx, st)
    n._lib.dwmh_set_stage_timing(n._ctx, 0)
    conv_ms, agg_ms, tc_ms, tc_tflop = float(st[0]), float(st[1]), float(st[2]), float(st[3])
    kinds_all = [n.layer_kernel_kind(i) for i in range(n.num_layers())]
    tc_launches_per_fwd = int(sum(kinds_all))
    # HBM-bound kernels timed alone
    def ev_time(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    zs_ms = ev_time(lambda: (flush.zero_(), vol.copy_(raw_dev[0]), n.normalize_(vol, None, 2))) - ev_time(lambda: (flush.zero_(), vol.copy_(raw_dev[0])))
    fin_ms = ev_time(lambda: (flush.zero_(), n.finalize(agg, wgt))) - ev_time(lambda: (flush.zero_(),))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    vps = world * args.steps / (ms_dev / 1e3)
    vps_e2e = world * max(1, args.steps) / (ms_e2e / 1e3)
    tflop_vol = flops_fwd * N_TILES * N_MIRRORS / 1e12
    n_batches = -(-N_TILES * N_MIRRORS // args.max_batch)
    tc_launches = tc_launches_per_fwd * n_batches
    achieved = tc_tflop / (tc_ms / 1e3)            # dominant kernel: conv3_tc_kernel (all tcgen05 conv / tconv launches)
    # bounded CPU sample: 3 tile-forwards of the same workload on all host cores
    cores = os.cpu_count()
    cpu_t = cpu_forward_sample(1 + args.cpu_forwards, cores)[1:] if (args.cpu_forwards > 0 and world == 1) else []      # N = 1 only
    cpu = None
    if cpu_t:
        per = sum(cpu_t) / len(cpu_t)
        cpu = {"value": 1.0 / (per * N_TILES * N_MIRRORS), "unit": "volumes/s", "cores": cores, "kind": "port",
               "sample": "%d of the 96 tile-forwards (128^3, softmax, x Gaussian/8) of the workload, fp32 oracle (torch CPU, %d threads), %.1f s per forward; volumes/s = 1/(96 x s per forward)" % (len(cpu_t), cores, per)}
    kinds = [n.layer_kernel_kind(i) for i in range(n.num_layers())]
    # DRAM traffic per conv3_tc_kernel launch: not measurable inside a timed run; taken from the committed ncu pass
    # (profiles/dram_traffic_r01.json, written by tools/dram_traffic.py from an ncu metrics run of the same 32-forward batch)
    ab, nl = conv_algorithmic_bytes(plans)
    alg_bytes = ab * args.max_batch / nl          # mean over the conv launches of one max_batch-forward batch
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic_r01.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["tc_dram_bytes_per_launch"], "profiles/dram_traffic_r01.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d conv3_tc_kernel launches of one 32-forward batch)" % tj["tc_launches"]
        except Exception:
            traffic = None
    out = {
        "metric": "volumes/sec", "value": vps, "unit": "volumes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if args.dtype == "fp16" else "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "forwards_per_volume": N_TILES * N_MIRRORS, "tflop_per_volume": tflop_vol,
                   "accumulate": "f32", "instance_norm_stats": "f64 sums", "batch_forwards": args.max_batch,
                   "parallelism": "cohort x%d (one subject stream per GPU, no collective)" % world,
                   "l2": "192 MiB flush buffer written between steps; per-batch activation working set (~7 GB) >> 126 MB L2",
                   "tcgen05_layers": int(sum(kinds)), "direct_layers": int(len(kinds) - sum(kinds))},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"],
                     "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel": "conv3_tc_kernel (tcgen05 implicit-GEMM conv3d / strided conv / transposed conv)",
                     "launches_per_volume": tc_launches, "avg_launch_ms": tc_ms / max(tc_launches, 1), "tflop_per_volume_in_kernel": tc_tflop,
                     "kernel_ms_per_volume": tc_ms, "kernel_share_of_step": tc_ms / (ms_dev / args.steps),
                     "conv_stack_ms_per_volume": conv_ms, "conv_stack_tflops": tflop_vol / (conv_ms / 1e3), "aggregate_ms_per_volume": agg_ms,
                     "peak_source": pk["src"] + ": sustained bf16 cuBLAS (the kernel runs inside a long step); burst %.1f" % pk["tensor_burst"]},
        "hbm_kernels": {"zscore": {"ms": zs_ms, "GBps": 12.0 * V / 1e9 / (zs_ms / 1e3), "frac_of_measured_hbm": 12.0 * V / 1e9 / (zs_ms / 1e3) / pk["hbm"]},
                        "finalize": {"ms": fin_ms, "GBps": 21.0 * V / 1e9 / (fin_ms / 1e3), "frac_of_measured_hbm": 21.0 * V / 1e9 / (fin_ms / 1e3) / pk["hbm"]},
                        "note": "volume (87 MB / 152 MB algorithmic) fits the 126 MB L2 only partly; timed with an L2 flush before each launch"},
        "cpu_baseline": cpu,
        "e2e": {"value": vps_e2e, "unit": "volumes/s", "h2d_bytes_per_step": 4 * V, "d2h_bytes_per_step": 9 * V,
                "ms_per_step": ms_e2e / max(1, args.steps), "api": "dwmh_predict_volume_host (C ABI, pinned host buffers)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--cpu-forwards", type=int, default=3, help="tile-forwards of the bounded CPU-baseline sample (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            # convenience: self-launch one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_b200(args)


if __name__ == "__main__":
    main()
