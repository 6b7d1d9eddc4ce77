"""CPU emulation of the CUDA path's numerics (SURVEY.md section 7 step 0 / risk R1).

Walks the oracle network layer by layer with the rounding points of the planned kernels:
operands of every tensor-core conv rounded to `op` (bf16 or fp16), fp32 accumulate, no conv
bias (cancels under InstanceNorm), fp32 instance statistics taken from the fp32 accumulator,
raw conv output stored as `store`, normalise+affine+LeakyReLU applied on the stored value and
re-rounded to `op` for the consumer.  Reports the parity-gate numbers of one tile against the
fp32 oracle.  Test infrastructure: imports oracle/.

usage: python tools/precision_probe.py [--op bf16|fp16] [--store bf16|fp16|fp32] [--mirrors 1|8]
"""
import argparse
import sys
import os
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402

DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


def rnd(x, dt):
    return x if dt == torch.float32 else x.to(dt).to(torch.float32)


def emu_block(blk, x, op, store, first=False):
    w = blk.conv.weight if first else rnd(blk.conv.weight, op)
    xin = x if first else rnd(x, op)
    y = F.conv3d(xin, w, None, blk.conv.stride, blk.conv.padding)
    mean = y.mean(dim=(2, 3, 4), keepdim=True)
    var = y.var(dim=(2, 3, 4), keepdim=True, unbiased=False)
    a = blk.instnorm.weight.view(1, -1, 1, 1, 1) / torch.sqrt(var + 1e-5)
    b = blk.instnorm.bias.view(1, -1, 1, 1, 1) - mean * a
    ys = rnd(y, store)
    return F.leaky_relu(a * ys + b, 0.01)


def emu_forward(net, x, op, store):
    skips = []
    first = True
    for d in range(len(net.conv_blocks_context) - 1):
        for blk in net.conv_blocks_context[d].blocks:
            x = emu_block(blk, x, op, store, first)
            first = False
        skips.append(x)
    for st in net.conv_blocks_context[-1]:
        for blk in st.blocks:
            x = emu_block(blk, x, op, store)
    for u in range(len(net.tu)):
        up = F.conv_transpose3d(rnd(x, op), rnd(net.tu[u].weight, op), None, net.tu[u].stride)
        x = torch.cat((rnd(up, op), skips[-(u + 1)]), 1)
        for st in net.conv_blocks_localization[u]:
            for blk in st.blocks:
                x = emu_block(blk, x, op, store)
    return net.seg_outputs[-1](x)      # head in fp32 on the un-rounded normalised value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="bf16")
    ap.add_argument("--store", default="bf16")
    ap.add_argument("--mirrors", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    net = O.build_benchmark_network(0)
    vol = O.synthetic_flair(seed=args.seed)
    vol[0] = O.zscore_nnunet(vol[0], np.where(vol[0] != 0, 0, -1), True)
    tile = torch.from_numpy(vol[:, 27:155, 45:173, 27:155].copy())[None]
    ref = torch.zeros(1, 2, 128, 128, 128)
    emu = torch.zeros(1, 2, 128, 128, 128)
    t0 = time.time()
    with torch.no_grad():
        for m in range(args.mirrors):
            dims = O.MIRROR_DIMS[m]
            xin = torch.flip(tile, dims) if dims else tile
            r = F.softmax(net(xin), 1)
            e = F.softmax(emu_forward(net, xin, DT[args.op], DT[args.store]), 1)
            if dims:
                r, e = torch.flip(r, dims), torch.flip(e, dims)
            ref += r / args.mirrors
            emu += e / args.mirrors
    d = (ref - emu).abs()[0, 1].numpy()
    sr, se = ref[0].argmax(0).numpy(), emu[0].argmax(0).numpy()
    print(f"op={args.op} store={args.store} mirrors={args.mirrors} time={time.time() - t0:.1f}s")
    print(f"  softmax |d| max={d.max():.4e} mean={d.mean():.4e} p99.9={np.quantile(d, 0.999):.4e}")
    print(f"  argmax agree={np.mean(sr == se):.6f} dice={O.hard_dice_binary(sr, se):.6f} fg={np.mean(sr > 0):.4f}")
    p1 = ref[0, 1].numpy()
    print(f"  frac |p-0.5|<0.01: {np.mean(np.abs(p1 - 0.5) < 0.01):.5f}")


if __name__ == "__main__":
    main()
