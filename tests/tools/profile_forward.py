"""Run warm forwards of the benchmark network on n patches (for `ncu` launch lists / captures).
usage: python tools/profile_forward.py [n_samples] [repeats] [generic]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402   (random-init weights only)
import deepwmh_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
plans = deepwmh_b200.benchmark_plans()
net = O.build_benchmark_network(0, plans)
tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=n)
tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
if len(sys.argv) > 3 and sys.argv[3] == "generic":
    tr.network.set_force_generic(True)
x = torch.randn(n, 1, 128, 128, 128, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(reps):
    tr.network.forward_patches(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); tr.network.forward_patches(x); b.record(); torch.cuda.synchronize()
print("forward of %d patches: %.2f ms" % (n, a.elapsed_time(b)))
