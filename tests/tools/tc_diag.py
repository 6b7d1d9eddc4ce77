"""GPU diagnostic: run the network once with the CUDA-core kernels forced and once with the tcgen05
kernels, and print per-layer deviations with a breakdown by channel / tile position / plane, so a wrong
descriptor or barrier protocol shows its pattern.  Test infrastructure (imports oracle/).

usage: python tools/tc_diag.py [small|bench] [n_samples]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
import deepwmh_b200  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    if which == "small":
        plans = deepwmh_b200.benchmark_plans(patch_size=(32, 32, 32), num_pool=3)
    elif which == "mid":
        plans = deepwmh_b200.benchmark_plans(patch_size=(64, 64, 64), num_pool=4)
    else:
        plans = deepwmh_b200.benchmark_plans()
    ps = tuple(int(i) for i in plans["plans_per_stage"][0]["patch_size"])
    net = O.build_benchmark_network(0, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=max(n, 2))
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    nw = tr.network
    x = torch.randn(n, 1, *ps, generator=torch.Generator().manual_seed(0)).cuda()
    L = nw.num_layers()
    kinds = [nw.layer_kernel_kind(i) for i in range(L)]
    print("layers:", L, "tcgen05:", sum(kinds), kinds)
    nw.set_force_generic(True)
    t0 = time.time(); p_ref = nw.forward_patches(x); torch.cuda.synchronize(); t_gen = time.time() - t0
    ref = [nw.layer_output(i, n).clone() for i in range(L)]
    nw.set_force_generic(False)
    t0 = time.time(); p_tc = nw.forward_patches(x); torch.cuda.synchronize(); t_tc = time.time() - t0
    got = [nw.layer_output(i, n).clone() for i in range(L)]
    print("forward generic %.1f ms, tcgen05 %.1f ms (first call each)" % (t_gen * 1e3, t_tc * 1e3))
    bad = None
    for i in range(L):
        d = (got[i] - ref[i]).abs()
        scale = ref[i].abs().max().item() + 1e-9
        rel = d.max().item() / scale
        flag = "TC" if kinds[i] else "  "
        print("layer %2d %s shape %-28s max|d|/max|ref| = %.3e  mean|d| = %.3e" % (i, flag, tuple(ref[i].shape), rel, d.mean().item()))
        if kinds[i] and rel > 1e-2 and bad is None:
            bad = i
    print("softmax max|d| tc vs generic: %.3e" % (p_tc - p_ref).abs().max().item())
    if bad is not None:
        i = bad
        d = (got[i] - ref[i]).abs()
        N, C, D, H, W = d.shape
        print("first bad tcgen05 layer %d: nan=%d inf=%d" % (i, torch.isnan(got[i]).sum().item(), torch.isinf(got[i]).sum().item()))
        print(" by sample  :", d.amax(dim=(1, 2, 3, 4)).cpu().numpy())
        print(" by channel :", np.array2string(d.amax(dim=(0, 2, 3, 4)).cpu().numpy(), precision=2, max_line_width=200))
        print(" by plane d :", np.array2string(d.amax(dim=(0, 1, 3, 4)).cpu().numpy(), precision=2, max_line_width=200))
        print(" by h       :", np.array2string(d.amax(dim=(0, 1, 2, 4)).cpu().numpy(), precision=2, max_line_width=200))
        print(" by w       :", np.array2string(d.amax(dim=(0, 1, 2, 3)).cpu().numpy(), precision=2, max_line_width=200))
        g, r = got[i][0, 0, D // 2], ref[i][0, 0, D // 2]
        print(" got[0,0,D/2,:8,:8]:\n", g[:8, :8].cpu().numpy())
        print(" ref[0,0,D/2,:8,:8]:\n", r[:8, :8].cpu().numpy())
    # timing of a warmed forward
    for force in (True, False):
        nw.set_force_generic(force)
        nw.forward_patches(x); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nw.forward_patches(x); b.record(); torch.cuda.synchronize()
        print("warm forward (%d samples) %s: %.2f ms" % (n, "generic" if force else "tcgen05", a.elapsed_time(b)))
    nw.set_force_generic(False)


if __name__ == "__main__":
    main()
