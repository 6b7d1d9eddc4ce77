"""CPU (this container, no GPU): where does the fp16 scheme lose argmax agreement?  Torch emulation of the library's rounding
points on single tiles of the benchmark volume against the fp32 oracle.  Test infrastructure (imports oracle/).

usage: python tests/tools/precision_cpu.py [variant ...]
  A variant is a comma list of switches on top of the library's default scheme (fp16 operands everywhere, fp16 raw storage
  above 64^3, fp32 raw storage at <= 64^3, first conv fp32-exact):
    raw32=<edge>   raw outputs with edge <= <edge> kept fp32          (default 64)
    act32=<edge>   activations NOT rounded to fp16 as conv operands in layers whose OUTPUT edge <= <edge>  (hi+lo split)
    w32=<edge>     weights NOT rounded in those layers
    tu32           transposed-conv operands / outputs not rounded
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402

H = torch.float16


def rnd(x, on=True):
    return x.to(H).to(torch.float32) if on else x


def parse(v):
    cfg = {"raw32": 64, "act32": 0, "w32": 0, "tu32": False}
    for tok in v.split(","):
        if not tok or tok == "base":
            continue
        if "=" in tok:
            k, val = tok.split("=")
            cfg[k] = int(val)
        else:
            cfg[tok] = True
    return cfg


def block(blk, x, first, cfg):
    edge_in = x.shape[-1]
    stride = blk.conv.stride[0]
    edge = edge_in // stride
    w = blk.conv.weight if (first or edge <= cfg["w32"]) else rnd(blk.conv.weight)
    xin = x if (first or edge <= cfg["act32"]) else rnd(x)
    y = F.conv3d(xin, w, None, blk.conv.stride, blk.conv.padding)
    mean = y.mean(dim=(2, 3, 4), keepdim=True)
    var = y.var(dim=(2, 3, 4), keepdim=True, unbiased=False)
    a = blk.instnorm.weight.view(1, -1, 1, 1, 1) / torch.sqrt(var + 1e-5)
    b = blk.instnorm.bias.view(1, -1, 1, 1, 1) - mean * a
    ys = y if edge <= cfg["raw32"] else rnd(y)
    return F.leaky_relu(a * ys + b, 0.01)


def emu_forward(net, x, cfg):
    skips = []
    first = True
    for d in range(len(net.conv_blocks_context) - 1):
        for blk in net.conv_blocks_context[d].blocks:
            x = block(blk, x, first, cfg); first = False
        skips.append(x)
    for st in net.conv_blocks_context[-1]:
        for blk in st.blocks:
            x = block(blk, x, False, cfg)
    for u in range(len(net.tu)):
        if cfg["tu32"]:
            up = F.conv_transpose3d(x, net.tu[u].weight, None, net.tu[u].stride)
        else:
            up = rnd(F.conv_transpose3d(rnd(x), rnd(net.tu[u].weight), None, net.tu[u].stride))
        x = torch.cat((up, skips[-(u + 1)]), 1)
        for st in net.conv_blocks_localization[u]:
            for blk in st.blocks:
                x = block(blk, x, False, cfg)
    return net.seg_outputs[-1](x)


def main():
    variants = sys.argv[1:] or ["base"]
    torch.set_num_threads(os.cpu_count())
    net = O.build_benchmark_network(0)
    raw = O.synthetic_flair(seed=0)
    data = raw.copy()
    data[0] = O.zscore_nnunet(raw[0], np.where(raw[0] != 0, 0, -1), True)
    tiles = [(0, 0, 0), (54, 90, 54), (0, 45, 54)]
    mirrors = [0, 7]
    refs = {}
    with torch.no_grad():
        for t in tiles:
            x = torch.from_numpy(np.ascontiguousarray(data[None, :, t[0]:t[0] + 128, t[1]:t[1] + 128, t[2]:t[2] + 128]))
            for m in mirrors:
                dims = O.MIRROR_DIMS[m]
                xin = torch.flip(x, dims) if dims else x
                t0 = time.time()
                refs[(t, m)] = (xin, torch.softmax(net(xin), 1))
        print("reference done (%.1f s per forward)" % (time.time() - t0), flush=True)
        for v in variants:
            cfg = parse(v)
            flips, tot, dmax = 0, 0, 0.0
            for (t, m), (xin, pr) in refs.items():
                p = torch.softmax(emu_forward(net, xin, cfg), 1)
                flips += int((p.argmax(1) != pr.argmax(1)).sum())
                tot += pr[0, 0].numel()
                dmax = max(dmax, float((p - pr).abs().max()))
            print("%-28s flips %7d of %d  agree %.6f  softmax|d| %.2e" % (v, flips, tot, 1 - flips / tot, dmax), flush=True)


if __name__ == "__main__":
    main()
