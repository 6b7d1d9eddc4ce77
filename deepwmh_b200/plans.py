"""`plans`-shaped dicts (the keys nnUNetTrainerV2.process_plans reads from plans.pkl)."""
from __future__ import annotations

from typing import Dict

import numpy as np


def benchmark_plans(patch_size=(128, 128, 128), num_pool: int = 5, base_num_features: int = 32) -> Dict:
    """The canonical 3d_fullres instance BASELINE.json is quoted on:
    1 modality, 2 classes, 5 stride-2 poolings, 3x3x3 kernels, patch 128^3."""
    return {
        "num_modalities": 1, "num_classes": 1,
        "base_num_features": base_num_features,
        "transpose_forward": [0, 1, 2], "transpose_backward": [0, 1, 2],
        "normalization_schemes": {0: "nonCT"}, "use_mask_for_norm": {0: True},
        "plans_per_stage": {0: {
            "patch_size": np.array(patch_size),
            "pool_op_kernel_sizes": [[2, 2, 2]] * num_pool,
            "conv_kernel_sizes": [[3, 3, 3]] * (num_pool + 1),
            "current_spacing": np.array([1.0, 1.0, 1.0]),
            "batch_size": 2,
        }},
    }
