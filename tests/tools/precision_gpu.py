"""GPU: where does the library's error come from?  Compares, on one 128^3 tile of the benchmark
volume, (a) the library, (b) the fp16 rounding-point emulation of tools/precision_probe.py, both
against (c) the fp32 oracle run with cuDNN (TF32 off) -- per layer for mirror 0 and for the 8-mirror
TTA softmax.  Test infrastructure (imports oracle/).

usage: python tools/precision_gpu.py [mirrors]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle as O  # noqa: E402
import deepwmh_b200  # noqa: E402
from precision_probe import emu_block, rnd  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def emu_forward_layers(net, x, op, store):
    outs, skips, first = [], [], True
    for d in range(len(net.conv_blocks_context) - 1):
        for blk in net.conv_blocks_context[d].blocks:
            x = emu_block(blk, x, op, store, first); first = False; outs.append(x)
        skips.append(x)
    for st in net.conv_blocks_context[-1]:
        for blk in st.blocks:
            x = emu_block(blk, x, op, store); outs.append(x)
    for u in range(len(net.tu)):
        up = F.conv_transpose3d(rnd(x, op), rnd(net.tu[u].weight, op), None, net.tu[u].stride)
        outs.append(rnd(up, op))
        x = torch.cat((rnd(up, op), skips[-(u + 1)]), 1)
        for st in net.conv_blocks_localization[u]:
            for blk in st.blocks:
                x = emu_block(blk, x, op, store); outs.append(x)
    return net.seg_outputs[-1](x), outs


def main():
    mirrors = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    plans = deepwmh_b200.benchmark_plans()
    net = O.build_benchmark_network(0, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=8)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    net = net.cuda()
    vol = O.synthetic_flair(seed=0)
    vol[0] = O.zscore_nnunet(vol[0], np.where(vol[0] != 0, 0, -1), True)
    tile = torch.from_numpy(vol[:, 27:155, 45:173, 27:155].copy())[None].cuda()
    xs = torch.cat([torch.flip(tile, O.MIRROR_DIMS[m]) if O.MIRROR_DIMS[m] else tile for m in range(mirrors)], 0).contiguous()
    # --- per-layer, mirror 0 ---
    acts = []
    hooks = [m.register_forward_hook(lambda mod, i, o: acts.append(o.detach()))
             for m in net.modules() if isinstance(m, (O.ConvDropoutNormNonlin, torch.nn.ConvTranspose3d))]
    with torch.no_grad():
        net(tile)
    for h in hooks:
        h.remove()
    with torch.no_grad():
        _, emu_acts = emu_forward_layers(net, tile, torch.float16, torch.float16)
    tr.network.forward_patches(xs[:1].contiguous())
    print("layer  shape                         lib_rms/ref_rms  emu_rms/ref_rms   ratio")
    for li, ref in enumerate(acts):
        lib = tr.network.layer_output(li, 1)
        rr = ref.float().pow(2).mean().sqrt().item()
        e_lib = (lib - ref).pow(2).mean().sqrt().item() / rr
        e_emu = (emu_acts[li] - ref).pow(2).mean().sqrt().item() / rr
        print("%3d    %-28s  %.3e        %.3e        %.2f" % (li, tuple(ref.shape), e_lib, e_emu, e_lib / max(e_emu, 1e-12)))
    # --- TTA softmax ---
    ref_p = torch.zeros(1, 2, 128, 128, 128, device="cuda"); emu_p = torch.zeros_like(ref_p); lib_p = torch.zeros_like(ref_p)
    lib_all = tr.network.forward_patches(xs)
    with torch.no_grad():
        for m in range(mirrors):
            dims = O.MIRROR_DIMS[m]
            r = F.softmax(net(xs[m:m + 1]), 1)
            e = F.softmax(emu_forward_layers(net, xs[m:m + 1], torch.float16, torch.float16)[0], 1)
            l = lib_all[m:m + 1]
            if dims:
                r, e, l = torch.flip(r, dims), torch.flip(e, dims), torch.flip(l, dims)
            ref_p += r / mirrors; emu_p += e / mirrors; lib_p += l / mirrors
    for name, p in (("lib", lib_p), ("emu", emu_p)):
        d = (p - ref_p).abs()[0, 1]
        agree = (p[0].argmax(0) == ref_p[0].argmax(0)).float().mean().item()
        print("%s vs fp32 oracle (%d mirrors): softmax |d| max %.3e mean %.3e  argmax agree %.6f" % (name, mirrors, d.max().item(), d.mean().item(), agree))
    d = (lib_p - emu_p).abs()[0, 1]
    print("lib vs emu: max %.3e mean %.3e" % (d.max().item(), d.mean().item()))


if __name__ == "__main__":
    main()
