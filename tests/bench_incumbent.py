"""Not a test (no test_ prefix): times the incumbent Blackwell path of SURVEY.md section 8d -- the PyTorch oracle
network on the same B200 through cuDNN (fp32, TF32, fp16 autocast = what nnU-Net's mixed_precision=True runs, channels-last
variant) -- next to this repo's forward on the same patches.  usage: python tests/bench_incumbent.py [n_patches]
Output is kept in profiles/incumbent_cudnn_r01.txt."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
import deepwmh_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
plans = deepwmh_b200.benchmark_plans()
net = O.build_benchmark_network(0, plans).cuda().eval()
net.do_ds = False
x = torch.randn(n, 1, 128, 128, 128, generator=torch.Generator().manual_seed(0)).cuda()
gflop = O.forward_flops(plans) / 1e9


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def fwd(autocast_dtype=None, channels_last=False):
    def f():
        with torch.no_grad():
            xi = x.contiguous(memory_format=torch.channels_last_3d) if channels_last else x
            if autocast_dtype is None:
                return torch.softmax(net(xi), 1)
            with torch.autocast("cuda", dtype=autocast_dtype):
                return torch.softmax(net(xi).float(), 1)
    return f


rows = []
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
rows.append(("cuDNN fp32 (no TF32)", timed(fwd())))
torch.backends.cudnn.allow_tf32 = True
rows.append(("cuDNN TF32", timed(fwd())))
rows.append(("cuDNN fp16 autocast (nnU-Net mixed_precision)", timed(fwd(torch.float16))))
rows.append(("cuDNN bf16 autocast", timed(fwd(torch.bfloat16))))
net_cl = net.to(memory_format=torch.channels_last_3d)
rows.append(("cuDNN fp16 autocast, channels_last_3d", timed(fwd(torch.float16, True))))
del net_cl
tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=n)
tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
rows.append(("this repo (tcgen05, fp16 in / fp32 acc)", timed(lambda: tr.network.forward_patches(x))))
print("%d patches of 128^3 per call, %.1f GFLOP per patch, %s" % (n, gflop, torch.cuda.get_device_name(0)))
for name, ms in rows:
    print("%-50s %9.2f ms  %8.2f ms/patch  %7.1f TFLOP/s" % (name, ms, ms / n, n * gflop / ms))
