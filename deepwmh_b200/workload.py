"""The synthetic workload of BASELINE.json / SURVEY.md section 8d, as product-side input generation for `bench.py` and the
command line's self-test: a synthetic T2-FLAIR-like volume, random-init Generic_UNet weights under their nnU-Net state-dict
names, and the algorithmic FLOP count of one patch forward.  Pure numpy; nothing here is on the measured path and nothing
imports `oracle/` (tests/test_workload.py holds these definitions to the oracle's)."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAX_NUM_FILTERS_3D = 320


def synthetic_flair(shape: Sequence[int] = (182, 218, 182), seed: int = 0) -> np.ndarray:
    """fp32 [1, X, Y, Z]: an ellipsoid head (semi-axes 0.45 x shape) filled with clip(N(100, 25), 1, .) tissue and 40 Gaussian
    hyper-intense blobs (sigma 1-4 voxels, +60..+150); exactly 0 outside the head, so the non-zero mask is the head."""
    shape = tuple(int(s) for s in shape)
    rng = np.random.default_rng(seed)
    grid = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    centre = [(s - 1) / 2.0 for s in shape]
    semi = [0.45 * s for s in shape]
    inside = sum(((grid[a] - centre[a]) / semi[a]) ** 2 for a in range(3)) <= 1.0
    vol = np.clip(rng.normal(100.0, 25.0, size=shape), 1.0, None).astype(np.float32)
    for _ in range(40):
        c = [rng.uniform(0.25 * s, 0.75 * s) for s in shape]
        sigma = rng.uniform(1.0, 4.0)
        amp = rng.uniform(60.0, 150.0)
        reach = int(math.ceil(4 * sigma))
        box = tuple(slice(max(0, int(c[a]) - reach), min(shape[a], int(c[a]) + reach + 1)) for a in range(3))
        d2 = sum((grid[a][box] - c[a]) ** 2 for a in range(3))
        vol[box] += (amp * np.exp(-d2 / (2 * sigma * sigma))).astype(np.float32)
    vol[~inside] = 0.0
    return vol[None].astype(np.float32)


def _stage(plans: Dict) -> Dict:
    return plans["plans_per_stage"][max(plans["plans_per_stage"].keys())]


def _features(plans: Dict) -> List[int]:
    feats = [int(plans["base_num_features"])]
    for _ in _stage(plans)["pool_op_kernel_sizes"]:
        feats.append(min(int(np.round(feats[-1] * 2)), MAX_NUM_FILTERS_3D))
    return feats


def forward_flops(plans: Dict) -> float:
    """2 x MAC of the convolutions, transposed convolutions and the last 1x1x1 head of one patch forward
    (954.46 GFLOP for the 128^3 benchmark instance)."""
    st = _stage(plans)
    pools, kers, feats = st["pool_op_kernel_sizes"], st["conv_kernel_sizes"], _features(plans)
    shape = np.array(st["patch_size"], dtype=np.int64)
    mac, cin, shapes = 0, int(plans["num_modalities"]), []
    for d in range(len(pools) + 1):
        if d > 0:
            shape = shape // np.array(pools[d - 1])
        mac += int(np.prod(shape)) * int(np.prod(kers[d])) * (cin * feats[d] + feats[d] * feats[d])
        cin = feats[d]
        shapes.append(shape.copy())
    for u in range(len(pools)):
        skip, sp = feats[-(2 + u)], shapes[-(2 + u)]
        mac += int(np.prod(sp)) * cin * skip
        mac += int(np.prod(sp)) * int(np.prod(kers[-(u + 1)])) * (2 * skip * skip + skip * skip)
        cin = skip
    mac += int(np.prod(shapes[0])) * cin * (int(plans["num_classes"]) + 1)
    return 2.0 * mac


def state_dict_layout(plans: Dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of every tensor of the nnU-Net Generic_UNet state dict for `plans` (state-dict naming)."""
    st = _stage(plans)
    pools, kers, feats = st["pool_op_kernel_sizes"], st["conv_kernel_sizes"], _features(plans)
    npool, ncls = len(pools), int(plans["num_classes"]) + 1
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def block(prefix: str, cin: int, cout: int, k: Sequence[int]):
        out.append((prefix + ".conv.weight", (cout, cin) + tuple(int(v) for v in k)))
        out.append((prefix + ".conv.bias", (cout,)))
        out.append((prefix + ".instnorm.weight", (cout,)))
        out.append((prefix + ".instnorm.bias", (cout,)))

    cin = int(plans["num_modalities"])
    for d in range(npool):
        block(f"conv_blocks_context.{d}.blocks.0", cin, feats[d], kers[d])
        block(f"conv_blocks_context.{d}.blocks.1", feats[d], feats[d], kers[d])
        cin = feats[d]
    block(f"conv_blocks_context.{npool}.0.blocks.0", cin, feats[npool], kers[npool])
    block(f"conv_blocks_context.{npool}.1.blocks.0", feats[npool], feats[npool], kers[npool])
    cur = feats[npool]
    for u in range(npool):
        skip = feats[-(2 + u)]
        block(f"conv_blocks_localization.{u}.0.blocks.0", 2 * skip, skip, kers[-(u + 1)])
        block(f"conv_blocks_localization.{u}.1.blocks.0", skip, skip, kers[-(u + 1)])
        out.append((f"tu.{u}.weight", (cur, skip) + tuple(int(v) for v in pools[-(u + 1)])))
        out.append((f"seg_outputs.{u}.weight", (ncls, skip, 1, 1, 1)))
        cur = skip
    return out


def random_init_state_dict(plans: Dict, model_index: int = 0) -> Dict[str, np.ndarray]:
    """Random-init weights of that architecture: InitWeights_He(1e-2) statistics for the (transposed) convolutions
    (normal, std = sqrt(2 / ((1 + 0.01^2) fan_in))), InstanceNorm gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1), conv bias ~ N(0, 0.1)
    so the affine and bias-cancellation paths carry signal.  Seeded by 1234 + model_index (k = 0..4 for the 5-model ensemble)."""
    rng = np.random.default_rng(1234 + int(model_index))
    sd: Dict[str, np.ndarray] = {}
    for key, shape in state_dict_layout(plans):
        if key.endswith("instnorm.weight"):
            v = rng.uniform(0.5, 1.5, size=shape)
        elif key.endswith(".bias"):
            v = rng.normal(0.0, 0.1, size=shape)
        else:
            # torch's fan_in = size(1) x receptive field, for Conv3d and ConvTranspose3d weights alike
            fan_in = shape[1] * int(np.prod(shape[2:]))
            v = rng.normal(0.0, math.sqrt(2.0 / ((1.0 + 0.01 ** 2) * fan_in)), size=shape)
        sd[key] = v.astype(np.float32)
    return sd
