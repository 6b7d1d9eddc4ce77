"""Spacing resample: the host-side decisions against the oracle restatement of
nnU-Net's resample_patient (CPU), the device kernels against scipy's spline arithmetic (GPU), and the file-level
drop-ins (`nnUNet_predict`-shaped entry, DeepWMH_predict) on a 0.9 x 0.9 x 3 mm case end to end (GPU).
The oracle of this row is parity-unpinned (nnunet and scikit-image are absent; see oracle/resample_oracle.py)."""
import os
import pickle

import numpy as np
import pytest
import torch

import oracle as O
from oracle import resample_oracle as R
from conftest import small_plans
from deepwmh_b200 import cli, nifti, preprocess

SPACINGS = [((1.0, 1.0, 1.0), (1.0, 1.0, 1.0)), ((3.0, 0.9, 0.9), (3.0, 1.0, 1.0)), ((3.0, 0.9, 0.9), (1.0, 1.0, 1.0)),
            ((0.9, 0.9, 3.0), (1.0, 1.0, 1.0)), ((1.2, 1.2, 1.2), (1.0, 1.0, 1.0)), ((0.24, 1.25, 1.25), (1.0, 1.0, 1.0)),
            ((5.0, 0.5, 0.5), (5.0, 0.43, 0.43)), ((1.0, 1.0, 1.0), (6.0, 0.9, 0.9)), ((2.0, 2.0, 2.0), (2.0, 2.0, 2.0))]


# ---------------------------------------------------------------------------------------------------------------------
# CPU: host logic and the oracle's known answers
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("orig,target", SPACINGS)
def test_shape_and_separate_z_rule_match_the_oracle(orig, target):
    for shape in [(20, 64, 64), (33, 57, 41), (7, 255, 256)]:
        assert np.array_equal(preprocess.resampled_shape(shape, orig, target), R.resampled_shape(shape, orig, target))
    for force in (None, True, False):
        assert preprocess.separate_z_rule(orig, target, force) == R.separate_z_rule(orig, target, force)


def test_separate_z_rule_known_answers():
    assert R.separate_z_rule((3.0, 0.9, 0.9), (1, 1, 1)) == (True, 0)                 # thick slices: anisotropy 3.33 > 3
    assert R.separate_z_rule((2.7, 0.9, 0.9), (1, 1, 1)) == (False, None)             # exactly 3 is not "> 3"
    assert R.separate_z_rule((1, 1, 1), (6.0, 0.9, 0.9)) == (True, 0)                 # target decides when the original is isotropic
    assert R.separate_z_rule((0.24, 1.25, 1.25), (1, 1, 1)) == (False, None)          # two tied low-res axes: switched off
    assert R.separate_z_rule((0.9, 3.0, 0.9), (1, 1, 1)) == (True, 1)
    assert np.array_equal(R.resampled_shape((20, 256, 256), (3.0, 0.9, 0.9), (1, 1, 1)), [60, 230, 230])
    assert np.array_equal(R.resampled_shape((5, 5, 5), (1.5, 2.5, 0.5), (1, 1, 1)), [8, 12, 2])   # numpy rounds half to even


def test_oracle_resize_properties():
    rng = np.random.default_rng(0)
    img = rng.normal(size=(9, 12, 10))
    assert np.array_equal(R.skimage_resize(img, img.shape, 3), img)
    const = np.full((6, 7, 8), 2.5)
    for order in (0, 1, 3):
        assert np.allclose(R.skimage_resize(const, (11, 5, 13), order), 2.5)
    ramp = np.arange(8, dtype=float)[:, None, None] * np.ones((8, 4, 4))
    up = R.skimage_resize(ramp, (16, 4, 4), 1)                       # half-pixel-centre grid, edge clamped
    assert np.allclose(up[1:-1, 0, 0], np.arange(1, 15) * 0.5 - 0.25) and up[0, 0, 0] == 0 and up[-1, 0, 0] == 7
    big = R.skimage_resize(img, (18, 24, 20), 3)
    assert big.min() >= img.min() and big.max() <= img.max()        # clip=True
    seg = np.where(rng.random((9, 12, 10)) > 0.4, 0, -1).astype(np.int8)
    rs = R.resize_segmentation(seg, (13, 17, 9), 1)
    ind = R.skimage_resize((seg == 0).astype(float), (13, 17, 9), 1)
    assert set(np.unique(rs)) <= {-1, 0} and np.array_equal(rs == 0, ind >= 0.5)


def test_oracle_separate_z_is_slicewise():
    rng = np.random.default_rng(1)
    data = rng.normal(size=(1, 5, 12, 10)).astype(np.float32)
    out = R.resample_data_or_seg(data, (15, 18, 7), False, 0, 3, True, 0)
    assert out.shape == (1, 15, 18, 7) and out.dtype == np.float32
    for o in range(15):                                               # nearest slice along the coarse axis
        src = int(np.floor(np.clip((o + 0.5) * 5 / 15 - 0.5, 0, 4) + 0.5))
        assert np.array_equal(out[0, o], R.skimage_resize(data[0, src], (18, 7), 3).astype(np.float32))


def test_console_scripts_resolve():
    import importlib
    import tomllib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scripts = tomllib.load(open(os.path.join(root, "pyproject.toml"), "rb"))["project"]["scripts"]
    assert set(scripts) >= {"DeepWMH_predict", "nnUNet_predict"}             # /root/reference/setup.py:47-55; predict.py:153
    for target in scripts.values():
        mod, fn = target.split(":")
        assert callable(getattr(importlib.import_module(mod), fn))


def test_nnunet_predict_rejects_what_it_does_not_serve(tmp_path, monkeypatch):
    from deepwmh_b200 import nnunet_predict as NP
    monkeypatch.setenv("RESULTS_FOLDER", str(tmp_path))
    base = ["-i", str(tmp_path), "-o", str(tmp_path / "o"), "-t", "Task002_FinalModel"]
    with pytest.raises(NotImplementedError):
        NP.main(base + ["-m", "2d"])
    with pytest.raises(NotImplementedError):
        NP.main(base + ["-f", "0", "1"])
    with pytest.raises(RuntimeError, match="Cannot find"):
        NP.main(base + ["-tr", "nnUNetTrainerV2", "-m", "3d_fullres", "-p", "nnUNetPlansv2.1", "-f", "all", "-chk", "model_best"])
    monkeypatch.delenv("RESULTS_FOLDER")
    with pytest.raises(RuntimeError, match="RESULTS_FOLDER"):
        NP.main(base)


# ---------------------------------------------------------------------------------------------------------------------
# GPU: kernels against scipy
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape,new_shape,order,sep", [
    ((20, 48, 40), (60, 43, 36), 3, None), ((20, 48, 40), (60, 43, 36), 3, 0), ((20, 48, 40), (20, 43, 36), 3, 0),
    ((33, 17, 29), (21, 40, 29), 3, None), ((12, 30, 9), (12, 41, 27), 3, 2), ((30, 12, 25), (44, 36, 31), 3, 1),
    ((20, 48, 40), (60, 43, 36), 1, None), ((20, 48, 40), (60, 43, 36), 1, 0), ((43, 36, 60), (48, 40, 20), 1, 2),
    ((5, 6, 7), (13, 11, 9), 3, None), ((64, 64, 64), (70, 58, 64), 3, None), ((9, 9, 9), (4, 5, 3), 1, None),
])
def test_device_resample_matches_scipy(shape, new_shape, order, sep):
    rng = np.random.default_rng(sum(shape) + order)
    vol = (rng.normal(100, 30, size=shape) * (rng.random(shape) > 0.2)).astype(np.float32)
    ref = R.resample_data_or_seg(vol[None], new_shape, False, sep, order, sep is not None, 0)[0]
    got = preprocess.resample_device(torch.from_numpy(vol).cuda(), new_shape, order, sep).cpu().numpy()
    assert got.shape == tuple(new_shape) and got.dtype == np.float32
    # fp64 arithmetic on both sides; the device prefilter is a 49-tap FIR of the same impulse response (|z|^24 = 2e-14)
    scale = float(np.abs(vol).max())
    assert np.abs(got - ref).max() <= 2e-6 * scale, np.abs(got - ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,new_shape,sep", [((20, 48, 40), (60, 43, 36), None), ((20, 48, 40), (60, 43, 36), 0),
                                                   ((31, 33, 35), (62, 66, 70), None), ((40, 40, 12), (35, 35, 30), 2)])
def test_device_crop_mask_resample_is_bit_exact(shape, new_shape, sep):
    """resize_segmentation(seg, order 1) of nnU-Net's {-1, 0} crop mask, incl. the `>= 0.5` ties that factor-2 zooms hit."""
    rng = np.random.default_rng(sum(new_shape))
    blob = rng.random(shape) > 0.5
    from scipy.ndimage import binary_dilation
    seg = np.where(binary_dilation(blob, iterations=2), 0, -1).astype(np.int8)
    ref = R.resample_data_or_seg(seg[None], new_shape, True, sep, 1, sep is not None, 0)[0]
    ind = torch.from_numpy((seg == 0).astype(np.float32)).cuda()
    got = preprocess.resample_device(ind, new_shape, 1, sep, out_mode=1).cpu().numpy()
    assert got.dtype == np.int8 and np.array_equal(got, ref)


# ---------------------------------------------------------------------------------------------------------------------
# GPU: a thick-slice case through the file-level entry points
# ---------------------------------------------------------------------------------------------------------------------
def _make_model_dir(root, plans, net, task="Task002_FinalModel", chk="model_best"):
    tdir = os.path.join(root, "nnUNet", "3d_fullres", task, "nnUNetTrainerV2__nnUNetPlansv2.1")
    os.makedirs(os.path.join(tdir, "all"), exist_ok=True)
    pickle.dump(plans, open(os.path.join(tdir, "plans.pkl"), "wb"))
    torch.save({"epoch": 1, "state_dict": net.state_dict()}, os.path.join(tdir, "all", chk + ".model"))
    return tdir


def _oracle_case(plans, net, vol_zyx, spacing_zyx, do_mirroring=True):
    """The oracle pipeline of predict_cases for one file: crop -> transpose -> resample_and_normalize -> tiled prediction
    -> transpose back -> resample back (order 1) -> argmax -> paste back."""
    tf, tb = plans["transpose_forward"], plans["transpose_backward"]
    cropped, seg, bbox = preprocess.crop_to_nonzero(vol_zyx[None])
    cropped = cropped.transpose([0] + [i + 1 for i in tf]); seg = seg.transpose([0] + [i + 1 for i in tf])
    target = np.array(plans["plans_per_stage"][0]["current_spacing"])
    data, _ = R.resample_and_normalize(cropped, seg, np.array(spacing_zyx)[tf], target, True)
    _, sm = O.OracleTrainer(plans, net).predict_preprocessed_data_return_seg_and_softmax(data.astype(np.float32), do_mirroring=do_mirroring)
    sm = sm.transpose([0] + [i + 1 for i in tb])
    sm = R.resample_softmax_back(sm, cropped.transpose([0] + [i + 1 for i in tb]).shape[1:], np.array(spacing_zyx), target)
    full = preprocess.paste_back(sm.argmax(0).astype(np.uint8), vol_zyx.shape, bbox)
    bg = np.ones(vol_zyx.shape, np.float32)
    bg[tuple(slice(b[0], b[1]) for b in bbox)] = sm[0]
    return full, bg


@pytest.mark.gpu
@pytest.mark.parametrize("target,tf", [((3.0, 1.0, 1.0), [0, 1, 2]), ((1.0, 1.0, 1.0), [0, 1, 2]), ((1.0, 3.0, 1.0), [1, 0, 2])])
def test_thick_slice_case_end_to_end(tmp_path, monkeypatch, target, tf):
    """0.9 x 0.9 x 3 mm FLAIR (NIfTI x, y, z) = spacing (3, 0.9, 0.9) in nnU-Net's z, y, x order: separate-z resample in,
    order-1 resample back out, through `nnUNet_predict ... --save_softmax --disable_tta` and through DeepWMH_predict."""
    from deepwmh_b200 import nnunet_predict as NP
    plans = small_plans()
    plans["plans_per_stage"][0]["current_spacing"] = np.array(target)
    plans["transpose_forward"] = tf
    plans["transpose_backward"] = [int(i) for i in np.argsort(tf)]
    net = O.build_benchmark_network(0, plans)
    root = str(tmp_path / "model")
    _make_model_dir(root, plans, net)
    vol_zyx = np.zeros((16, 60, 56), np.float32)
    vol_zyx[1:15, 4:57, 3:52] = O.synthetic_flair((14, 53, 49), seed=21)[0]
    vol_zyx[1:15, 4:57, 3:52][vol_zyx[1:15, 4:57, 3:52] == 0] = 2.0
    indir = tmp_path / "in"; indir.mkdir()
    nifti.write_nifti(str(indir / "caseA_0000.nii.gz"), np.transpose(vol_zyx, (2, 1, 0)), nifti.default_header((56, 60, 16), spacing=(0.9, 0.9, 3.0)))
    monkeypatch.setenv("RESULTS_FOLDER", root)
    out = str(tmp_path / "out")
    argv = ["-i", str(indir), "-o", out, "-tr", "nnUNetTrainerV2", "-m", "3d_fullres", "-p", "nnUNetPlansv2.1", "-t", "Task002_FinalModel",
            "-f", "all", "-chk", "model_best", "--disable_post_processing", "--selected_cases", "caseA"]
    assert NP.main(argv + ["--save_softmax", "--disable_tta"]) == 0
    assert os.path.isfile(os.path.join(out, "plans.pkl"))
    seg_xyz, hdr = nifti.read_nifti(os.path.join(out, "caseA.nii.gz"))
    bg_xyz, _ = nifti.read_nifti(os.path.join(out, "caseA_0.nii.gz"))
    assert np.allclose(hdr["spacing"], (0.9, 0.9, 3.0)) and seg_xyz.shape == (56, 60, 16)
    full_ref, bg_ref = _oracle_case(plans, net, vol_zyx, (3.0, 0.9, 0.9), do_mirroring=False)
    seg = np.transpose(seg_xyz, (2, 1, 0)); bg = np.transpose(bg_xyz, (2, 1, 0))
    assert np.mean(seg == full_ref) > 0.997, np.mean(seg == full_ref)
    assert np.abs(bg - bg_ref).max() < 2e-2 and (bg[0] == 1.0).all()
    # a second run finds its outputs and skips; DeepWMH_predict (TTA on) produces the output tree on the same case
    assert NP.main(argv + ["--save_softmax", "--disable_tta"]) == 0
    out2 = str(tmp_path / "out2")
    assert cli.main(["-i", str(indir / "caseA_0000.nii.gz"), "-n", "caseA", "-m", root, "-o", out2, "--skip-bfc", "-g", "0"]) == 0
    seg2, _ = nifti.read_nifti(os.path.join(out2, "002_Segmentations", "001_raw", "caseA.nii.gz"))
    full_tta, _ = _oracle_case(plans, net, vol_zyx, (3.0, 0.9, 0.9), do_mirroring=True)
    assert np.mean(np.transpose(seg2, (2, 1, 0)) == full_tta) > 0.997
    assert os.path.isfile(os.path.join(out2, "002_Segmentations", "002_postproc_3mm", "caseA.nii.gz"))
