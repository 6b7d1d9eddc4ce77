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
@pytest.mark.parametrize("base,patch,nb", [(32, (72, 88, 44), 2), (64, (72, 88, 44), 2), (32, (72, 80, 40), 1)])
def test_norm_on_load_layers_match_separate_norm_pass(base, patch, nb):
    """Layers whose InstanceNorm + LeakyReLU are applied by the consumer's loader warps (raw fp16 producer, one stride-1
    consumer) against the CUDA-core path with its separate norm pass: partial tiles in both in-plane axes (88 = 5.5 x 16,
    44 = 5.5 x 8: the loaders' zero-once border slots), 32-channel sources (one chunk step per plane) and 64-channel sources
    (two steps per plane, coefficients re-read per chunk); one sample of 72 x 80 x 40 gives 25 tiles x 15 z-blocks, an odd item
    count, i.e. the single-group variant of the kernel (8 warps, its own role table)."""
    import deepwmh_b200
    plans = small_plans(patch=patch, pools=((2, 2, 2),) * 2, base=base)
    ps = patch
    net = O.build_benchmark_network(0, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=nb)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    nw = tr.network
    x = torch.randn(nb, 1, *ps, generator=torch.Generator().manual_seed(3)).cuda()
    L = nw.num_layers()
    fused = [i for i in range(L) if nw.layer_norm_on_load(i) == 1]
    assert fused, "no layer of this plan is normalised on load"
    nw.set_force_generic(True)
    assert all(nw.layer_norm_on_load(i) == 0 for i in range(L))
    p_ref = nw.forward_patches(x).clone()
    ref = [nw.layer_output(i, nb).clone() for i in range(L)]
    nw.set_force_generic(False)
    p_tc = nw.forward_patches(x).clone()
    for i in range(L):
        got = nw.layer_output(i, nb)
        assert torch.isfinite(got).all(), i
        rel = (got - ref[i]).abs().max().item() / (ref[i].abs().max().item() + 1e-9)
        assert rel < 1e-2, (base, patch, i, i in fused, rel)
    assert (p_tc - p_ref).abs().max().item() < 5e-3
    if all(nw.layer_kernel_kind(i) == 1 for i in range(L)):          # (a 64-channel first conv runs on the CUDA-core kernel: fp64 atomics)
        assert torch.equal(nw.forward_patches(x), p_tc)              # deterministic
    with torch.no_grad():
        p_or = torch.softmax(net(x.cpu()), 1)
    assert (p_tc.cpu() - p_or).abs().max().item() < 1e-2
    nw.close()
