"""One stage-1 case (182x218x182, k = 10 references) for ncu: python tests/tools/stage1_profile.py  (see profiles/)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepwmh_b200 import stage1 as S  # noqa: E402

shape, K = (182, 218, 182), 10
rng = np.random.default_rng(0)
g = np.stack(np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij"))
brain = torch.from_numpy(((g ** 2).sum(0) < 0.8).astype(np.float32)).cuda()
vols = [torch.from_numpy(((100 + rng.normal(0, 10, shape)) * brain.cpu().numpy()).astype(np.float32)).cuda() for _ in range(K + 1)]
for _ in range(2):                                   # the second pass is the one to read
    r = S.nll_anomaly_map(vols[0], vols[1:], brain, brain, intensity_prior="+")
    m = S.median_3mm(r["anomaly"], [1.0, 1.0, 1.0])
    m2 = S.median_filter(r["anomaly"], [4, 4, 4])
torch.cuda.synchronize()
print("ok", float(m.sum()))
