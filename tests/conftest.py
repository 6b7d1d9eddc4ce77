import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def small_plans(patch=(32, 32, 32), pools=((2, 2, 2),) * 3, kernels=None, base=32):
    kernels = kernels if kernels is not None else [[3, 3, 3]] * (len(pools) + 1)
    return {
        "num_modalities": 1, "num_classes": 1, "base_num_features": base,
        "transpose_forward": [0, 1, 2], "transpose_backward": [0, 1, 2],
        "normalization_schemes": {0: "nonCT"}, "use_mask_for_norm": {0: True},
        "plans_per_stage": {0: {"patch_size": np.array(patch), "pool_op_kernel_sizes": [list(p) for p in pools],
                                "conv_kernel_sizes": [list(k) for k in kernels],
                                "current_spacing": np.array([1.0, 1.0, 1.0]), "batch_size": 2}},
    }


@pytest.fixture(scope="session")
def plans_small():
    return small_plans()


@pytest.fixture(scope="session")
def plans_aniso():
    # anisotropic 2-D-FLAIR-like plan: first stage does not pool / convolve along x
    return small_plans(patch=(16, 64, 48), pools=((1, 2, 2), (2, 2, 2), (2, 2, 2)),
                       kernels=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]])
