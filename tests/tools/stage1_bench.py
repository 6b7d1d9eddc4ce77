"""Device timing of the stage-1 kernels at BASELINE.json's 182x218x182 volume, k = 10 references
(what the reference registers per case), against the measured HBM peak; the numpy / scipy restatement
(oracle/intree_oracle.py) is timed beside it on a bounded sample.  usage: python tests/tools/stage1_bench.py [--cpu]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepwmh_b200 import stage1 as S  # noqa: E402


def timed(fn, iters=10, flush=None):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)                                           # 256 MiB write: evicts the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    peak = 6538.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbps"])
    except Exception:
        pass
    shape, K = (182, 218, 182), 10
    V = int(np.prod(shape))
    rng = np.random.default_rng(0)
    g = np.stack(np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij"))
    brain_h = ((g ** 2).sum(0) < 0.8).astype(np.float32)
    tgt_h = ((100 + rng.normal(0, 10, shape)) * brain_h).astype(np.float32)
    refs_h = [((100 + rng.normal(0, 10, shape)) * brain_h).astype(np.float32) for _ in range(K)]
    brain = torch.from_numpy(brain_h).cuda()
    valid = brain.clone()
    tgt = torch.from_numpy(tgt_h).cuda()
    refs = [torch.from_numpy(r).cuda() for r in refs_h]
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
    patch = S.image_patch_size([1.0, 1.0, 1.0])
    zt = S.z_score(tgt, brain, fill_outside=True)
    zr = [S.z_score(r, brain, fill_outside=True) for r in refs]
    an = S.nll(zt, zr, min_std=0.03, side="+", mul_mask=valid)
    rows = []

    def row(name, fn, alg_bytes):
        ms = timed(fn, flush=flush)
        rows.append({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(alg_bytes / 1e6, 1),
                     "GBps": round(alg_bytes / ms / 1e6, 1), "frac_of_measured_hbm": round(alg_bytes / ms / 1e6 / peak, 3)})

    # z_score: stats pass reads x + mask, apply pass reads x + mask and writes x (the wrapper's defensive copy is extra)
    row("z_score + tissue-min fill (2 kernels + copy)", lambda: S.z_score(tgt, brain, fill_outside=True), V * 4 * (2 + 2 + 1 + 2))
    row("mean_std_grid (3 kernels)", lambda: S.mean_std_grid(zt, patch, mask=valid), V * 4 * (2 + 2))
    row("group mean/std + nll, k=%d" % K, lambda: S.nll(zt, zr, min_std=0.03, side="+", return_all=True, mul_mask=valid), V * 4 * (K + 2 + 3))
    row("median 3x3x3", lambda: S.median_3mm(an, [1.0, 1.0, 1.0]), V * 4 * 2)
    row("median 4x4x4", lambda: S.median_filter(an, [4, 4, 4]), V * 4 * 2)
    row("median 6x6 slices", lambda: S.median_filter(an, [6, 6, 1]), V * 4 * 2)
    row("median 6x5x3 (generic rank search)", lambda: S.median_filter(an, [6, 5, 3]), V * 4 * 2)
    row("component_filtering (3 orientations, 16 launches)", lambda: S.component_filtering(valid, [1.0, 1.0, 1.0]), V * 4 * 2)
    row("anomaly map end to end (k=%d)" % K, lambda: S.nll_anomaly_map(tgt, refs, brain, valid, intensity_prior="+"), V * 4 * (K + 1) * 2)
    # the whole nll_analysis on arrays (masks, anomaly map, component filtering, K reference maps, histogram curves and
    # threshold, label vote, priors, median): wall clock with a synchronise, since it contains small host steps
    lab1 = [brain] * K
    t_h = np.zeros(shape, np.float32); t_h[brain_h > 0.5] = 3; t_h[(g ** 2).sum(0) < 0.6] = 1; t_h[((g ** 2).sum(0) < 0.7) & (g[2] < -0.4)] = 2
    lab2 = [torch.from_numpy(t_h).cuda()] * K
    whole = lambda: S.nll_analysis_arrays(tgt, refs, lab1, lab2, [1.0, 1.0, 1.0], apply_otsu=True, intensity_prior="+")
    whole(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.time(); whole(); torch.cuda.synchronize(); ts.append(time.time() - t0)
    rows.append({"kernel": "nll_analysis on arrays, whole function (k=%d, apply_otsu)" % K, "ms": round(1e3 * float(np.median(ts)), 3),
                 "note": "wall clock incl. host steps (256- / 400-bin scans, pointer tables) and small D2H syncs"})
    out = {"shape": shape, "k_refs": K, "hbm_peak_GBps": peak, "l2": "256 MiB buffer written before every timed call", "rows": rows}
    if "--cpu" in sys.argv:
        from oracle import intree_oracle as I
        t0 = time.time()
        I.nll_anomaly_arrays(tgt_h, refs_h[:3], brain_h, brain_h, patch)
        t1 = time.time()
        I.median_3mm(an.cpu().numpy(), [1.0, 1.0, 1.0])
        t2 = time.time()
        I.component_filtering(brain_h, [1.0, 1.0, 1.0])
        t3 = time.time()
        I.nll_analysis_arrays(tgt_h, refs_h[:3], [brain_h] * 3, [t_h] * 3, [1.0, 1.0, 1.0], "+", apply_otsu=True)
        t4 = time.time()
        out["cpu_port"] = {"anomaly_map_k3_s": round(t1 - t0, 2), "median_3x3x3_s": round(t2 - t1, 2),
                           "component_filtering_s": round(t3 - t2, 2), "nll_analysis_arrays_k3_s": round(t4 - t3, 2), "cores": os.cpu_count(),
                           "note": "numpy / scipy restatement (oracle/intree_oracle.py), k = 3 of the 10 references"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
