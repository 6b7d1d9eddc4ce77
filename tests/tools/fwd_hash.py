"""GPU: SHA-1 of the softmax of n seeded patches through the benchmark network (bitwise A/B of kernel variants / knobs).
usage: fwd_hash.py [n]"""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deepwmh_b200  # noqa: E402
from deepwmh_b200 import workload as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
plans = deepwmh_b200.benchmark_plans()
tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=n)
tr.load_checkpoint_ram({"state_dict": W.random_init_state_dict(plans, 0)}, False)
x = torch.randn(n, 1, 128, 128, 128, generator=torch.Generator().manual_seed(7)).cuda()
hs = []
for _ in range(3):
    y = tr.network.forward_patches(x)
    torch.cuda.synchronize()
    hs.append(hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:16])
print("softmax sha1 x3:", hs, "mean %.6f" % float(y.mean()))
