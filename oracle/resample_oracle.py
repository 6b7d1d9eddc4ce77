"""TEST INFRASTRUCTURE ONLY -- CPU restatement of nnU-Net v1's spacing resampling, the step
`nnUNet_predict` runs on every case whose voxel spacing differs from the plans' target spacing (reached from
deepwmh/main/predict.py:153-156 and deepwmh/pipeline/DCNN_multistage.py:331-344,531-535).

PARITY UNPINNED.  The arithmetic lives in two third-party packages that are neither vendored under /root/reference nor
importable here: `nnunet` (fork github.com/lchdl/nnUNet_for_DeepWMH of MIC-DKFZ/nnUNet v1, no pinned version,
README.md:69-80) -- `preprocessing/preprocessing.py::{get_do_separate_z, get_lowres_axis, resample_patient,
resample_data_or_seg}`, `inference/segmentation_export.py::save_segmentation_nifti_from_softmax` -- and scikit-image
(`setup.py:32-46` lists it without a version) -- `skimage.transform.resize`.  Both are restated from their published
algorithms:
  * skimage.transform.resize(img, shape, order, mode='edge', anti_aliasing=False) (>= 0.19) is
    scipy.ndimage.zoom(img, shape_out / shape_in, order=order, mode='nearest', grid_mode=True, prefilter=True) followed by
    a clip to [img.min(), img.max()] (clip=True); scipy IS installed here, so the spline arithmetic itself is scipy's.
  * batchgenerators.augmentations.utils.resize_segmentation(seg, shape, order=1): per label, resize of the 0/1 mask and
    `>= 0.5`, labels processed in ascending order.
Only tests/ (and smoke / the bench's CPU legs) may import this module.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.ndimage import map_coordinates, zoom

RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD = 3


def skimage_resize(img: np.ndarray, out_shape: Sequence[int], order: int) -> np.ndarray:
    """skimage.transform.resize(img, out_shape, order, mode='edge', anti_aliasing=False, clip=True) for a float image."""
    img = np.asarray(img, dtype=np.float64)
    out_shape = tuple(int(v) for v in out_shape)
    if out_shape == img.shape:
        return img.copy()
    factors = [o / i for o, i in zip(out_shape, img.shape)]
    out = zoom(img, factors, order=order, mode="nearest", grid_mode=True, prefilter=True)
    assert out.shape == out_shape, (out.shape, out_shape)
    if order > 0:
        np.clip(out, img.min(), img.max(), out=out)
    return out


def get_do_separate_z(spacing, anisotropy_threshold=RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD) -> bool:
    """[U:preprocessing.py::get_do_separate_z]"""
    return bool((np.max(spacing) / np.min(spacing)) > anisotropy_threshold)


def get_lowres_axis(new_spacing) -> np.ndarray:
    """[U:preprocessing.py::get_lowres_axis]"""
    return np.where(max(new_spacing) / np.array(new_spacing) == 1)[0]


def resampled_shape(shape: Sequence[int], original_spacing, target_spacing) -> np.ndarray:
    """[U:preprocessing.py::resample_patient] new_shape = round(original_spacing / target_spacing * shape)."""
    return np.round((np.array(original_spacing) / np.array(target_spacing)).astype(float) * np.array(shape)).astype(int)


def separate_z_rule(original_spacing, target_spacing, force_separate_z=None,
                    threshold=RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD) -> Tuple[bool, Optional[int]]:
    """[U:preprocessing.py::resample_patient] the (do_separate_z, axis) decision, incl. the 'two equal low-res axes' escape."""
    if force_separate_z is not None:
        do_separate_z = bool(force_separate_z)
        axis = get_lowres_axis(original_spacing) if force_separate_z else None
    elif get_do_separate_z(original_spacing, threshold):
        do_separate_z, axis = True, get_lowres_axis(original_spacing)
    elif get_do_separate_z(target_spacing, threshold):
        do_separate_z, axis = True, get_lowres_axis(target_spacing)
    else:
        do_separate_z, axis = False, None
    if axis is not None:
        if len(axis) == 3 or len(axis) == 2:
            do_separate_z = False
    return do_separate_z, (int(axis[0]) if (do_separate_z and axis is not None) else None)


def resample_data_or_seg(data: np.ndarray, new_shape, is_seg: bool, axis: Optional[int] = None, order: int = 3,
                         do_separate_z: bool = False, order_z: int = 0) -> np.ndarray:
    """[U:preprocessing.py::resample_data_or_seg]  data [c, x, y, z]; separate-z: per-slice 2-D resize, then the low-res
    axis by map_coordinates(order_z, mode='nearest') on the half-pixel-centre grid."""
    assert data.ndim == 4
    resize_fn = resize_segmentation if is_seg else (lambda a, s, o: skimage_resize(a, s, o))
    dtype_data = data.dtype
    shape = np.array(data[0].shape)
    new_shape = np.array([int(v) for v in new_shape])
    if not np.any(shape != new_shape):
        return data
    data = data.astype(float)
    if do_separate_z:
        assert axis is not None
        new_shape_2d = [new_shape[i] for i in range(3) if i != axis]
        final = []
        for c in range(data.shape[0]):
            slices = []
            for sid in range(shape[axis]):
                sl = [slice(None)] * 3
                sl[axis] = sid
                slices.append(resize_fn(data[c][tuple(sl)], new_shape_2d, order))
            reshaped = np.stack(slices, axis)
            if shape[axis] != new_shape[axis]:
                rows, cols, dim = new_shape
                orig_rows, orig_cols, orig_dim = reshaped.shape
                map_rows, map_cols, map_dims = np.mgrid[:rows, :cols, :dim]
                map_rows = float(orig_rows) / rows * (map_rows + 0.5) - 0.5
                map_cols = float(orig_cols) / cols * (map_cols + 0.5) - 0.5
                map_dims = float(orig_dim) / dim * (map_dims + 0.5) - 0.5
                coord_map = np.array([map_rows, map_cols, map_dims])
                if not is_seg or order_z == 0:
                    final.append(map_coordinates(reshaped, coord_map, order=order_z, mode="nearest")[None])
                else:
                    labels = np.unique(reshaped)
                    out = np.zeros(new_shape, dtype=dtype_data)
                    for cl in labels:
                        m = np.round(map_coordinates((reshaped == cl).astype(float), coord_map, order=order_z, mode="nearest"))
                        out[m > 0.5] = cl
                    final.append(out[None])
            else:
                final.append(reshaped[None])
        return np.vstack(final).astype(dtype_data)
    return np.vstack([resize_fn(data[c], new_shape, order)[None] for c in range(data.shape[0])]).astype(dtype_data)


def resize_segmentation(segmentation: np.ndarray, new_shape, order: int = 3) -> np.ndarray:
    """[U:batchgenerators.augmentations.utils::resize_segmentation]"""
    tpe = segmentation.dtype
    new_shape = tuple(int(v) for v in new_shape)
    if order == 0:
        return skimage_resize(segmentation.astype(float), new_shape, 0).astype(tpe)
    reshaped = np.zeros(new_shape, dtype=segmentation.dtype)
    for c in np.unique(segmentation):
        multihot = skimage_resize((segmentation == c).astype(float), new_shape, order)
        reshaped[multihot >= 0.5] = c
    return reshaped


def resample_patient(data: Optional[np.ndarray], seg: Optional[np.ndarray], original_spacing, target_spacing, order_data=3,
                     order_seg=0, force_separate_z=False, order_z_data=0, order_z_seg=0,
                     separate_z_anisotropy_threshold=RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD):
    """[U:preprocessing.py::resample_patient]"""
    assert not (data is None and seg is None)
    shape = np.array((data if data is not None else seg)[0].shape)
    new_shape = resampled_shape(shape, original_spacing, target_spacing)
    do_separate_z, axis = separate_z_rule(original_spacing, target_spacing, force_separate_z, separate_z_anisotropy_threshold)
    d = resample_data_or_seg(data, new_shape, False, axis, order_data, do_separate_z, order_z_data) if data is not None else None
    s = resample_data_or_seg(seg, new_shape, True, axis, order_seg, do_separate_z, order_z_seg) if seg is not None else None
    return d, s


def resample_and_normalize(data: np.ndarray, seg: np.ndarray, original_spacing, target_spacing, use_mask_for_norm: bool,
                           force_separate_z=None):
    """[U:preprocessing.py::GenericPreprocessor.resample_and_normalize], single non-CT modality: data / seg already cropped
    and transposed; resample (data order 3, seg order 1, order_z 0), seg[seg < -1] = 0, masked z-score."""
    from .nnunet_oracle import zscore_nnunet
    data, seg = resample_patient(data, seg, np.array(original_spacing), np.array(target_spacing), 3, 1,
                                 force_separate_z=force_separate_z, order_z_data=0, order_z_seg=0)
    seg = seg.copy()
    seg[seg < -1] = 0
    out = data.copy()
    for c in range(data.shape[0]):
        out[c] = zscore_nnunet(data[c], seg[-1], use_mask_for_norm)
    return out, seg


def resample_softmax_back(softmax: np.ndarray, shape_after_cropping, original_spacing, spacing_after_resampling,
                          force_separate_z=None, interpolation_order=1, interpolation_order_z=0) -> np.ndarray:
    """[U:inference/segmentation_export.py::save_segmentation_nifti_from_softmax] the resample-back of the class
    probabilities [classes, x, y, z] (after transpose_backward) to the cropped original grid, before argmax."""
    if not np.any([i != j for i, j in zip(np.array(softmax.shape[1:]), np.array(shape_after_cropping))]):
        return softmax
    if force_separate_z is None:
        if get_do_separate_z(original_spacing):
            do_separate_z, lowres_axis = True, get_lowres_axis(original_spacing)
        elif get_do_separate_z(spacing_after_resampling):
            do_separate_z, lowres_axis = True, get_lowres_axis(spacing_after_resampling)
        else:
            do_separate_z, lowres_axis = False, None
    else:
        do_separate_z = bool(force_separate_z)
        lowres_axis = get_lowres_axis(original_spacing) if do_separate_z else None
    if lowres_axis is not None and len(lowres_axis) != 1:
        do_separate_z = False
    return resample_data_or_seg(softmax, shape_after_cropping, False, int(lowres_axis[0]) if do_separate_z else None,
                                interpolation_order, do_separate_z, interpolation_order_z)
