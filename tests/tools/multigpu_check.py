"""torchrun --nproc-per-node N tools/multigpu_check.py : the N>1 modes on real GPUs (NCCL).
 * tile-sharded prediction of one volume (contiguous tile ranges per rank + ONE all-reduce of agg/wgt) must equal the
   single-GPU result up to fp32 summation order / fp16 statistics noise;
 * cohort sharding: every rank predicts its own subjects, results gathered on the host.
Test infrastructure (imports oracle/ for synthetic inputs and weights)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
import deepwmh_b200  # noqa: E402
from deepwmh_b200.parallel import predict_volume_tile_sharded, shard_cohort  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = tuple(int(v) for v in os.environ.get("MG_SHAPE", "182,218,182").split(","))
plans = deepwmh_b200.benchmark_plans()
net = O.build_benchmark_network(0, plans)
tr = deepwmh_b200.nnUNetTrainerV2(plans, device=local, max_batch=32)
tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
raw = O.synthetic_flair(shape, seed=0)
data = raw.copy()
data[0] = O.zscore_nnunet(raw[0], np.where(raw[0] != 0, 0, -1), True)

# --- tile-sharded ---
for _ in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
    seg, probs = predict_volume_tile_sharded(tr, data)
    torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t0
if rank == 0:
    seg1, p1 = tr.predict_preprocessed_data_return_seg_and_softmax(data)
    d = np.abs(probs.cpu().numpy() - p1).max()
    agree = float(np.mean(seg.cpu().numpy() == seg1.astype(np.uint8)))
    print("tile-sharded x%d on %s: %.3f s/volume; vs single GPU: softmax max|d| %.2e, argmax agree %.6f" % (world, shape, dt, d, agree))
    assert d < 3e-3 and agree > 0.9995
# --- cohort ---
mine = shard_cohort(2 * world, rank, world)
dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
fg = []
for s in mine:
    v = O.synthetic_flair((182, 218, 182), seed=s)[0]
    sg, _ = tr.predict_raw_volume_host(v)
    fg.append((s, float(sg.mean())))
torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t0
out = [None] * world
dist.all_gather_object(out, fg)
if rank == 0:
    allfg = sorted(sum(out, []))
    assert [s for s, _ in allfg] == list(range(2 * world))
    print("cohort of %d subjects on %d GPUs (incl. host-side synthesis): %.2f s; fg fractions %s" % (2 * world, world, dt, ["%.3f" % f for _, f in allfg]))
dist.destroy_process_group()
