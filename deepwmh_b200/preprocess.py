"""File-level steps either side of the hot path, host side.

Restates, for the single-modality case DeepWMH uses, what nnU-Net v1 does around predict_3D:
  * crop_to_nonzero                      [U:preprocessing/cropping.py]      (scipy binary_fill_holes, as upstream)
  * paste-back export                    [U:inference/segmentation_export.py]
  * remove_3mm_sparks                    deepwmh/analysis/image_ops.py:325-367 (scipy.ndimage.label, as the reference)
Resampling to the plans' target spacing is NOT implemented: a volume whose spacing differs from
plans['plans_per_stage'][stage]['current_spacing'] by more than 1 % raises NotImplementedError.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy.ndimage import binary_fill_holes, label


def crop_to_nonzero(data: np.ndarray) -> Tuple[np.ndarray, np.ndarray, List[List[int]]]:
    """data [c, z, y, x] -> (cropped data, seg [1, z, y, x] with -1 outside the filled nonzero mask and 0 inside,
    bbox [[z0, z1], [y0, y1], [x0, x1]])."""
    assert data.ndim == 4
    mask = np.zeros(data.shape[1:], dtype=bool)
    for c in range(data.shape[0]):
        mask |= data[c] != 0
    mask = binary_fill_holes(mask)
    idx = np.where(mask)
    if idx[0].size == 0:
        bbox = [[0, s] for s in data.shape[1:]]
    else:
        bbox = [[int(idx[a].min()), int(idx[a].max()) + 1] for a in range(3)]
    sl = tuple(slice(b[0], b[1]) for b in bbox)
    cropped = data[(slice(None),) + sl]
    seg = np.where(mask[sl], 0, -1).astype(np.int8)[None]
    return np.ascontiguousarray(cropped), seg, bbox


def check_spacing(spacing: Sequence[float], target: Sequence[float]):
    s, t = np.asarray(spacing, dtype=np.float64), np.asarray(target, dtype=np.float64)
    if np.any(np.abs(s - t) > 0.01 * t):
        raise NotImplementedError("resampling is not implemented: image spacing %s differs from the plans' target "
                                  "spacing %s (SURVEY.md section 8f-1)" % (tuple(s), tuple(t)))


def paste_back(seg_cropped: np.ndarray, original_shape: Sequence[int], bbox: List[List[int]]) -> np.ndarray:
    out = np.zeros(tuple(original_shape), dtype=np.uint8)
    out[tuple(slice(b[0], b[1]) for b in bbox)] = seg_cropped
    return out


def remove_sparks(mask: np.ndarray, min_volume: int = 3) -> np.ndarray:
    """Discard connected components (6-connectivity, scipy default) smaller than min_volume voxels."""
    m = (mask > 0.5).astype(np.int32)
    lab, n = label(m)
    if n == 0:
        return np.zeros_like(m)
    sizes = np.bincount(lab.ravel(), minlength=n + 1)
    keep = sizes >= min_volume
    keep[0] = False
    return keep[lab].astype(np.int32)


def remove_3mm_sparks(mask: np.ndarray, voxel_size: Sequence[float]) -> np.ndarray:
    """deepwmh/analysis/image_ops.py:346-367: components below 3 mm^3 (>= 2 voxels) are removed; for thick-slice
    data (anisotropy > 3) the limit is 3 voxels."""
    vs = [float(v) for v in voxel_size]
    if max(vs) / min(vs) > 3.0:
        return remove_sparks(mask, 3)
    mv = int(np.around(3.0 / (vs[0] * vs[1] * vs[2])))
    return remove_sparks(mask, max(mv, 2))
