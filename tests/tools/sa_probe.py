"""GPU: forward time of one 32-patch batch (set DWMH_TC_MAX_SA / other knobs in the environment).  usage: sa_probe.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deepwmh_b200  # noqa: E402
from deepwmh_b200 import workload as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
plans = deepwmh_b200.benchmark_plans()
tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=n)
tr.load_checkpoint_ram({"state_dict": W.random_init_state_dict(plans, 0)}, False)
x = torch.randn(n, 1, 128, 128, 128, device="cuda")
for _ in range(2):
    tr.network.forward_patches(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    tr.network.forward_patches(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("DWMH_TC_MAX_SA=%s  %d patches: %.2f ms  (%.1f TFLOP/s)" % (os.environ.get("DWMH_TC_MAX_SA", "4"), n, ms, n * W.forward_flops(plans) / 1e9 / ms))
