"""Host-only behaviour of the dwmh_s1_* entry points (no GPU: only workspace sizing and argument validation, which run
before any CUDA call): error convention = non-zero return + dwmh_last_error()."""
import ctypes as C

import numpy as np
import pytest

from deepwmh_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def err(lib):
    return lib.dwmh_last_error().decode()


def i3(*v):
    return (C.c_int32 * 3)(*v)


def test_mean_std_grid_workspace_follows_the_reference_geometry(lib):
    n = C.c_int64(0)
    # 182x218x182, patch 50: step 25, padded 200x250x200 -> 8x10x8 cells, bordered grid 10x12x10
    assert lib.dwmh_s1_mean_std_grid_workspace(182, 218, 182, i3(50, 50, 50), C.byref(n)) == 0
    a256 = lambda v: (v + 255) // 256 * 256
    assert n.value == a256(8 * 10 * 8 * 3 * 8) + 2 * a256(10 * 12 * 10 * 8)
    # odd patch sizes are rounded up to even (2 * ceil(p / 2)): 9 -> 10, step 5
    assert lib.dwmh_s1_mean_std_grid_workspace(40, 46, 38, i3(9, 14, 11), C.byref(n)) == 0
    # patch -> (10, 14, 12), step (5, 7, 6), padded (40, 56, 48) -> 8 x 8 x 8 cells
    assert n.value == a256(8 * 8 * 8 * 3 * 8) + 2 * a256(10 * 10 * 10 * 8)
    m = C.c_int64(0)
    assert lib.dwmh_s1_local_mean_align_workspace(182, 218, 182, i3(50, 50, 50), 10, C.byref(m)) == 0
    assert m.value == a256(11 * 640 * 3 * 8) + 2 * 11 * a256(1200 * 8)
    assert lib.dwmh_s1_mean_std_grid_workspace(0, 4, 4, i3(2, 2, 2), C.byref(n)) != 0 and "empty volume" in err(lib)
    assert lib.dwmh_s1_mean_std_grid_workspace(4, 4, 4, i3(2, 0, 2), C.byref(n)) != 0 and "patch_size[1]" in err(lib)
    assert lib.dwmh_s1_local_mean_align_workspace(4, 4, 4, i3(2, 2, 2), 33, C.byref(m)) != 0 and "k = 33" in err(lib)


def test_component_filtering_workspace(lib):
    n = C.c_int64(0)
    assert lib.dwmh_s1_component_filtering_workspace(182, 218, 182, C.byref(n)) == 0
    V = 182 * 218 * 182
    a256 = lambda v: (v + 255) // 256 * 256
    assert n.value == 3 * a256(V * 4) + a256(218 * 8)
    assert lib.dwmh_s1_component_filtering_workspace(4, -1, 4, C.byref(n)) != 0 and "empty volume" in err(lib)


def test_argument_validation_happens_before_any_cuda_call(lib):
    one = C.c_void_p(16)                                              # a non-null dummy address; never dereferenced on these paths
    ptrs = (C.c_void_p * 1)(16)
    assert lib.dwmh_s1_zscore(0, None, None, 8, 0, one, None, None) != 0 and "null argument" in err(lib)
    assert lib.dwmh_s1_zscore(0, one, None, 0, 0, one, None, None) != 0 and "empty volume" in err(lib)
    assert lib.dwmh_s1_zscore(0, one, None, 8, 1, one, None, None) != 0 and "fill_outside needs a mask" in err(lib)
    assert lib.dwmh_s1_zscore_batch(0, ptrs, 34, None, 8, 0, one, None) != 0 and "34 volumes" in err(lib)
    assert lib.dwmh_s1_group_nll(0, one, ptrs, 0, 0.03, 0, None, one, None, None, 8, None) != 0 and "k = 0" in err(lib)
    assert lib.dwmh_s1_group_nll(0, one, ptrs, 1, 0.03, 2, None, one, None, None, 8, None) != 0 and "side" in err(lib)
    assert lib.dwmh_s1_group_nll_masked(0, one, ptrs, None, 1, 0.03, 0, None, one, None, None, 8, None) != 0 and "null argument" in err(lib)
    assert lib.dwmh_s1_median_filter(0, one, one, 4, 4, 4, i3(3, 3, 3), None) != 0 and "must not alias" in err(lib)
    assert lib.dwmh_s1_median_filter(0, one, C.c_void_p(32), 4, 4, 4, i3(3, 10, 3), None) != 0 and "1..9 per axis" in err(lib)
    assert lib.dwmh_s1_histogram(0, one, None, 8, 0, 0.0, one, 4096, one, None) != 0 and "nbins = 4096" in err(lib)
    assert lib.dwmh_s1_label_vote(0, ptrs, 1, 17, one, one, 8, None) != 0 and "17 label ids" in err(lib)
    assert lib.dwmh_s1_apply_priors(0, one, None, one, None, 2, 8, None) != 0 and "stage 2 needs" in err(lib)
    assert lib.dwmh_s1_apply_priors(0, one, None, one, None, 3, 8, None) != 0 and "stage must be 1 or 2" in err(lib)
    vs = (C.c_double * 3)(1.0, 0.0, 1.0)
    assert lib.dwmh_s1_component_filtering(0, one, 4, 4, 4, vs, C.c_void_p(32), one, None) != 0 and "voxel_size[1]" in err(lib)
    assert lib.dwmh_s1_masked_sums(0, ptrs, 0, None, 8, 0, one, (C.c_double * 3)(), None) != 0 and "0 volumes" in err(lib)
