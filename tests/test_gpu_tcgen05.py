"""The tcgen05 implicit-GEMM conv against the shape-complete CUDA-core kernels (same rounded weights,
same fp32 accumulation; only the summation order differs) layer by layer, plus SASS evidence that the
library really carries tcgen05 / TMA / TMEM instructions."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import oracle as O
from conftest import small_plans


def test_sass_has_blackwell_native_instructions():
    """CPU: cuobjdump of the in-tree library shows UTCHMMA (tcgen05.mma), UTMALDG / UBLKCP (TMA), LDTM (TMEM load), UTCBAR (tcgen05.commit)."""
    from deepwmh_b200 import build as b
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", b.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, mnemonic


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["small", "mid", "aniso", "base64", "deep"])
def test_tcgen05_layers_match_cuda_core_layers(cfg):
    import deepwmh_b200
    if cfg == "small":
        plans = small_plans()
    elif cfg == "mid":
        plans = small_plans(patch=(64, 48, 40), pools=((2, 2, 2),) * 3)      # partial tiles: H, W not multiples of 16 / 8
    elif cfg == "base64":                                                     # 64-channel first conv (CUDA-core fallback), wider layers
        plans = small_plans(patch=(32, 40, 24), pools=((2, 2, 2),) * 2, base=64)
    elif cfg == "deep":                                                       # 5 poolings: 320-channel streamed layers on 2^3 planes
        plans = small_plans(patch=(64, 64, 64), pools=((2, 2, 2),) * 5)
    else:
        plans = small_plans(patch=(16, 64, 48), pools=((1, 2, 2), (2, 2, 2), (2, 2, 2)),
                            kernels=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]])
    ps = tuple(int(i) for i in plans["plans_per_stage"][0]["patch_size"])
    net = O.build_benchmark_network(0, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=3)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    nw = tr.network
    x = torch.randn(3, 1, *ps, generator=torch.Generator().manual_seed(0)).cuda()
    L = nw.num_layers()
    kinds = [nw.layer_kernel_kind(i) for i in range(L)]
    assert sum(kinds) >= (6 if cfg != "base64" else 4), kinds     # every 3x3x3 conv with Cin % 16 == 0
    nw.set_force_generic(True)
    p_ref = nw.forward_patches(x)
    ref = [nw.layer_output(i, 3).clone() for i in range(L)]
    nw.set_force_generic(False)
    p_tc = nw.forward_patches(x)
    for i in range(L):
        got = nw.layer_output(i, 3)
        rel = (got - ref[i]).abs().max().item() / (ref[i].abs().max().item() + 1e-9)
        assert rel < 1e-2, (cfg, i, kinds[i], rel)
    assert (p_tc - p_ref).abs().max().item() < 5e-3
    with torch.no_grad():
        p_or = torch.softmax(net(x.cpu()), 1)
    assert (p_tc.cpu() - p_or).abs().max().item() < 1e-2
    tr.network.close()


@pytest.mark.gpu
def test_cluster_multicast_of_streamed_weights_matches_plain_launch():
    """DWMH_TC_CLUSTER=1: CTA pairs (thread-block clusters) receive every streamed weight tile by one TMA multicast.  Same
    arithmetic, so the layer outputs must agree with the plain launch (up to the order of the statistics atomics)."""
    import deepwmh_b200
    plans = small_plans(patch=(32, 40, 24), pools=((2, 2, 2),) * 2, base=64)      # 128-channel layers: streamed, 4 tiles per plane
    net = O.build_benchmark_network(0, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=3)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    nw = tr.network
    x = torch.randn(3, 1, 32, 40, 24, generator=torch.Generator().manual_seed(1)).cuda()
    old = os.environ.get("DWMH_TC_CLUSTER")
    try:
        os.environ["DWMH_TC_CLUSTER"] = "0"
        p0 = nw.forward_patches(x).clone()
        ref = [nw.layer_output(i, 3).clone() for i in range(nw.num_layers())]
        os.environ["DWMH_TC_CLUSTER"] = "1"
        p1 = nw.forward_patches(x)
        for i in range(nw.num_layers()):
            got = nw.layer_output(i, 3)
            assert torch.isfinite(got).all(), i
            assert (got - ref[i]).abs().max().item() <= 2e-3 * ref[i].abs().max().item() + 1e-6, i
        assert (p1 - p0).abs().max().item() < 2e-3
    finally:
        if old is None:
            os.environ.pop("DWMH_TC_CLUSTER", None)
        else:
            os.environ["DWMH_TC_CLUSTER"] = old
    nw.close()
