"""GPU parity tests proper: every call goes through the C ABI (libdeepwmh_b200.so) and is compared
with the CPU oracle on the same seeded inputs, with the committed golden fixture at BASELINE.json's
full size, and through size-independent properties.

Gate (BASELINE.json north_star): softmax |d| <= 2e-2, argmax agreement >= 99.9 %, Dice >= 0.999.
Operand/storage type is fp16 with fp32 accumulation (profiles/precision_probe_r01.txt: bf16 cannot
meet the argmax gate on random-init weights); tolerances below are stated per test.
"""
import os

import numpy as np
import pytest
import torch

import oracle as O
from conftest import small_plans

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SOFTMAX_TOL, AGREE_MIN, DICE_MIN = 2e-2, 0.999, 0.999


def _trainer(plans, model_index=0, act_dtype="fp16", max_batch=8):
    import deepwmh_b200
    net = O.build_benchmark_network(model_index, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, act_dtype=act_dtype, max_batch=max_batch)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    return tr, net


@pytest.fixture(scope="module")
def small():
    plans = small_plans()
    tr, net = _trainer(plans)
    yield tr, net, plans
    tr.network.close()


def _gate(seg_ref, p_ref, seg, p, tol=SOFTMAX_TOL, agree=AGREE_MIN, dice=DICE_MIN):
    rep = O.parity_report(seg_ref, p_ref, seg, p)
    assert rep["softmax_max_abs"] <= tol, rep
    assert rep["argmax_agree"] >= agree, rep
    if 0.02 < rep["fg_frac_ref"] < 0.98:
        assert rep["dice"] >= dice, rep
    return rep


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4096, 182 * 218 * 182, 1000003])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_zscore_matches_oracle(small, n, mode):
    tr = small[0]
    rng = np.random.default_rng(n + mode)
    v = np.clip(rng.normal(100, 25, size=n), 1, None).astype(np.float32)
    v[rng.random(n) < 0.3] = 0
    seg = np.where(v != 0, 0, -1).astype(np.int8)
    if mode == 0:
        ref = O.zscore_nnunet(v, None, False)
    else:
        ref = O.zscore_nnunet(v, seg, True)
    dv = torch.from_numpy(v).cuda()
    ds = torch.from_numpy(seg).cuda() if mode == 1 else None
    mean, std, cnt = tr.network.normalize_(dv, ds, mask_mode=mode)
    got = dv.cpu().numpy()
    assert cnt == (n if mode == 0 else int((v != 0).sum()))
    assert np.abs(got - ref).max() < 2e-5          # fp64 statistics vs numpy's fp32 pairwise sums
    if mode:
        assert (got[v == 0] == 0).all()


def test_forward_patches_match_oracle(small):
    tr, net, _ = small
    x = torch.randn(3, 1, 32, 32, 32, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = torch.softmax(net(x), 1).numpy()
    got = tr.network.forward_patches(x.cuda()).cpu().numpy()
    d = np.abs(got - ref)
    assert d.max() < 1e-2 and d.mean() < 1e-3, (d.max(), d.mean())
    assert np.mean(got.argmax(1) == ref.argmax(1)) > 0.998


def test_every_layer_matches_oracle(small):
    """Layer-level parity: normalised+activated output of each conv / transposed conv."""
    tr, net, _ = small
    x = torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(1))
    acts = []
    hooks = []
    for m in net.modules():
        if isinstance(m, (O.ConvDropoutNormNonlin, torch.nn.ConvTranspose3d)):
            hooks.append(m.register_forward_hook(lambda mod, i, o: acts.append(o.detach().clone())))
    with torch.no_grad():
        net(x)
    for h in hooks:
        h.remove()
    tr.network.forward_patches(x.cuda())
    assert tr.network.num_layers() == len(acts)
    for li, ref in enumerate(acts):
        got = tr.network.layer_output(li, n=2).cpu()
        assert got.shape == ref.shape, (li, got.shape, ref.shape)
        err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        assert err < 2e-2, (li, err)


def test_anisotropic_plan_matches_oracle():
    plans = small_plans(patch=(16, 64, 48), pools=((1, 2, 2), (2, 2, 2), (2, 2, 2)),
                        kernels=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]])
    tr, net = _trainer(plans)
    x = torch.randn(2, 1, 16, 64, 48, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = torch.softmax(net(x), 1).numpy()
    got = tr.network.forward_patches(x.cuda()).cpu().numpy()
    assert np.abs(got - ref).max() < 1e-2
    # the 1x3x3 / stride (1,2,2) layers run on the tensor cores too; only the 1-channel first conv (9 taps) stays direct
    kinds = [tr.network.layer_kernel_kind(i) for i in range(tr.network.num_layers())]
    assert kinds[0] == 0 and all(k == 1 for k in kinds[1:]), kinds
    tr.network.set_force_generic(True)                       # CUDA-core cross-check of the same plan
    got_g = tr.network.forward_patches(x.cuda()).cpu().numpy()
    tr.network.set_force_generic(False)
    assert np.abs(got_g - ref).max() < 1e-2 and np.abs(got_g - got).max() < 1e-2
    tr.network.close()


def test_anisotropic_plan_every_layer_matches_oracle():
    """Layer-level parity of a thick-slice plan: two 1x3x3 stages with (1,2,2) pooling, then 3x3x3."""
    plans = small_plans(patch=(8, 64, 64), pools=((1, 2, 2), (1, 2, 2), (2, 2, 2)),
                        kernels=[[1, 3, 3], [1, 3, 3], [3, 3, 3], [3, 3, 3]])
    tr, net = _trainer(plans)
    x = torch.randn(2, 1, 8, 64, 64, generator=torch.Generator().manual_seed(5))
    acts, hooks = [], []
    for m in net.modules():
        if isinstance(m, (O.ConvDropoutNormNonlin, torch.nn.ConvTranspose3d)):
            hooks.append(m.register_forward_hook(lambda mod, i, o: acts.append(o.detach().clone())))
    with torch.no_grad():
        net(x)
    for h in hooks:
        h.remove()
    tr.network.forward_patches(x.cuda())
    for li, ref in enumerate(acts):
        got = tr.network.layer_output(li, n=2).cpu()
        assert got.shape == ref.shape, (li, got.shape, ref.shape)
        err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        assert err < 2e-2, (li, err)
    tr.network.close()


@pytest.mark.parametrize("do_mirroring,mirror_axes,use_gaussian,shape", [
    (True, (0, 1, 2), True, (40, 50, 45)),
    (False, (0, 1, 2), True, (40, 50, 45)),
    (True, (0, 2), True, (33, 32, 47)),
    (True, (0, 1, 2), False, (40, 34, 32)),
    (False, (), True, (20, 32, 40)),            # smaller than the patch on axis 0: padded (a10)
    (True, (0, 1, 2), True, (32, 32, 32)),      # single tile: Gaussian not applied
])
def test_predict_matches_oracle(small, do_mirroring, mirror_axes, use_gaussian, shape):
    tr, net, plans = small
    data = O.synthetic_flair(shape, seed=3)
    data[0] = O.zscore_nnunet(data[0], np.where(data[0] != 0, 0, -1), True)
    otr = O.OracleTrainer(plans, net)
    seg_r, p_r = otr.predict_preprocessed_data_return_seg_and_softmax(data, do_mirroring, mirror_axes or None, True, 0.5, use_gaussian)
    seg, p = tr.predict_preprocessed_data_return_seg_and_softmax(data, do_mirroring, mirror_axes or None, True, 0.5, use_gaussian)
    assert seg.shape == seg_r.shape and seg.dtype == seg_r.dtype and p.shape == p_r.shape and p.dtype == np.float32
    # 50-90k voxels with ~10 % foreground: a dozen boundary flips already move Dice by 1e-3, so the
    # small-volume gate is 1e-2 softmax / 99.8 % / 0.995; the full BASELINE gate is applied at full size below
    _gate(seg_r, p_r, seg, p, tol=1e-2, agree=0.998, dice=0.995)
    assert np.allclose(p.sum(0), 1.0, atol=1e-5)


def test_unsupported_combinations_raise(small):
    tr = small[0]
    x = np.zeros((1, 32, 32, 32), np.float32)
    with pytest.raises(NotImplementedError):
        tr.network.predict_3D(x, True, use_sliding_window=False)
    with pytest.raises(NotImplementedError):
        tr.network.predict_3D(x, True, use_sliding_window=True, regions_class_order=(1,))
    with pytest.raises(AssertionError):
        tr.network.predict_3D(x, True, use_sliding_window=True, step_size=1.5)
    with pytest.raises(AssertionError):
        tr.network.predict_3D(x[0], True, use_sliding_window=True)


def test_host_buffer_call_equals_device_call_and_is_deterministic(small):
    tr, net, plans = small
    raw = O.synthetic_flair((40, 50, 45), seed=4)[0]
    seg_h, p_h = tr.predict_raw_volume_host(raw)
    seg_h2, p_h2 = tr.predict_raw_volume_host(raw)
    # no atomics on the path: ordered overlap-add, InstanceNorm statistics combined in a fixed order
    assert np.array_equal(p_h, p_h2) and np.array_equal(seg_h, seg_h2)
    vol = torch.from_numpy(raw.copy()).cuda()
    tr.network.normalize_(vol, None, 2)
    seg_d, p_d = tr.predict_preprocessed_data_return_seg_and_softmax(vol.cpu().numpy()[None])
    assert np.array_equal(p_h, p_d) and np.array_equal(seg_h, seg_d.astype(np.uint8))


def test_tile_ranges_sum_to_full_run(small):
    """Single-GPU rehearsal of the tile-sharded mode: two ranks' ranges + a sum == one run."""
    tr, net, _ = small
    from deepwmh_b200.parallel import shard_tiles
    data = O.synthetic_flair((40, 50, 45), seed=5)
    vol = torch.from_numpy(data[0]).cuda()
    n_tiles = tr.network.num_tiles(vol.shape, 0.5)
    bufs = []
    for rank in range(2):
        agg = torch.zeros((2,) + tuple(vol.shape), device="cuda"); wgt = torch.zeros(tuple(vol.shape), device="cuda")
        b, e = shard_tiles(n_tiles, rank, 2)
        tr.network.accumulate_tiles(vol, agg, wgt, 0.5, True, (0, 1, 2), True, b, e)
        bufs.append((agg, wgt))
    agg = torch.zeros((2,) + tuple(vol.shape), device="cuda"); wgt = torch.zeros(tuple(vol.shape), device="cuda")
    tr.network.accumulate_tiles(vol, agg, wgt, 0.5, True, (0, 1, 2), True)
    # the forwards are deterministic; the two partial buffers are summed in another fp32 order than the single run
    assert torch.allclose(bufs[0][1] + bufs[1][1], wgt, rtol=1e-6)
    assert ((bufs[0][0] + bufs[1][0]) / wgt - agg / wgt).abs().max().item() < 1e-5
    nb = torch.from_numpy(O.predict_3D_tiled(_ConstNet(), data, 0.5, False, (), (32, 32, 32), True, return_buffers=True)[1][0])
    assert torch.allclose(wgt.cpu(), nb, rtol=1e-6)


class _ConstNet:
    num_classes = 2
    _gaussian_3d = None
    _patch_size_for_gaussian_3d = None

    def inference_apply_nonlin(self, x):
        return x

    def __call__(self, x):
        return torch.zeros(x.shape[0], 2, *x.shape[2:])


def test_bf16_mode_runs_with_looser_agreement():
    plans = small_plans()
    tr, net = _trainer(plans, act_dtype="bf16")
    x = torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref = torch.softmax(net(x), 1).numpy()
    got = tr.network.forward_patches(x.cuda()).cpu().numpy()
    assert np.abs(got - ref).max() < 8e-2
    tr.network.close()


def test_ensemble_mean_of_two_models():
    plans = small_plans()
    data = O.synthetic_flair((40, 34, 32), seed=7)
    data[0] = O.zscore_nnunet(data[0], np.where(data[0] != 0, 0, -1), True)
    from deepwmh_b200.parallel import ensemble_mean
    ps, ps_ref = [], []
    for k in range(2):
        tr, net = _trainer(plans, model_index=k)
        ps.append(torch.from_numpy(tr.predict_preprocessed_data_return_seg_and_softmax(data)[1]))
        ps_ref.append(O.OracleTrainer(plans, net).predict_preprocessed_data_return_seg_and_softmax(data)[1])
        tr.network.close()
    got = ensemble_mean(ps).numpy()
    ref = np.mean(np.stack(ps_ref), 0)
    _gate(ref.argmax(0), ref, got.argmax(0), got, tol=1e-2, agree=0.998, dice=0.995)


def test_checkpoint_reload_in_one_context_matches_fresh_contexts():
    """a13 as predict_cases does it: ONE trainer, `for p in params: trainer.load_checkpoint_ram(p, False)`."""
    plans = small_plans()
    data = O.synthetic_flair((40, 34, 32), seed=8)
    data[0] = O.zscore_nnunet(data[0], np.where(data[0] != 0, 0, -1), True)
    nets = [O.build_benchmark_network(k, plans) for k in range(2)]
    tr, _ = _trainer(plans, model_index=0)
    outs = []
    for k in (0, 1, 0):
        tr.load_checkpoint_ram({"state_dict": nets[k].state_dict()}, False)
        outs.append(tr.predict_preprocessed_data_return_seg_and_softmax(data)[1])
    assert np.array_equal(outs[0], outs[2])                # same weights again -> bit-identical result
    assert np.abs(outs[0] - outs[1]).max() > 5e-2          # different model really loaded
    ref1 = O.OracleTrainer(plans, nets[1]).predict_preprocessed_data_return_seg_and_softmax(data)[1]
    assert np.abs(outs[1] - ref1).max() < 1e-2
    tr.network.close()


# ------------------------------------------------------------------------------------------------
# BASELINE.json full size (config 2): 182x218x182, 128^3 patch, 8x TTA, against the committed fixture
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full():
    import deepwmh_b200
    plans = deepwmh_b200.benchmark_plans()
    tr, net = _trainer(plans)
    yield tr
    tr.network.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "v1_tta.npz")), reason="golden fixture missing")
def test_full_size_tta_matches_golden(full):
    g = np.load(os.path.join(GOLD, "v1_tta.npz"))
    raw = O.synthetic_flair((182, 218, 182), seed=0)[0]
    seg, p = full.predict_raw_volume_host(raw)
    seg_ref = np.unpackbits(g["seg_bits"])[: seg.size].reshape(seg.shape)
    agree = float(np.mean(seg == seg_ref))
    dice = O.hard_dice_binary(seg_ref, seg)
    d_lat = np.abs(p[1, ::3, ::3, ::3] - g["p1_lattice"]).max()
    d_blk = max(np.abs(p[1, 60:108, 80:128, 60:108] - g["p1_block"]).max(), np.abs(p[0, 60:108, 80:128, 60:108] - g["p0_block"]).max())
    print("full-size parity: agree=%.6f dice=%.6f softmax|d| lattice=%.3e block=%.3e" % (agree, dice, d_lat, d_blk))
    assert max(d_lat, d_blk) <= SOFTMAX_TOL and agree >= AGREE_MIN and dice >= DICE_MIN


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "v1_notta.npz")), reason="golden fixture missing")
def test_full_size_no_tta_matches_golden(full):
    g = np.load(os.path.join(GOLD, "v1_notta.npz"))
    raw = O.synthetic_flair((182, 218, 182), seed=0)[0]
    seg, p = full.predict_raw_volume_host(raw, do_mirroring=False)
    seg_ref = np.unpackbits(g["seg_bits"])[: seg.size].reshape(seg.shape)
    agree = float(np.mean(seg == seg_ref))
    d_lat = np.abs(p[1, ::3, ::3, ::3] - g["p1_lattice"]).max()
    print("full-size no-TTA parity: agree=%.6f dice=%.6f softmax|d|=%.3e" % (agree, O.hard_dice_binary(seg_ref, seg), d_lat))
    assert d_lat <= SOFTMAX_TOL and agree >= 0.9985          # single pass, no TTA averaging (probe: 0.99917/tile)


def test_full_size_properties(full):
    """Size-independent properties at the full benchmark size."""
    raw = O.synthetic_flair((182, 218, 182), seed=1)[0]
    seg, p = full.predict_raw_volume_host(raw, do_mirroring=False)
    assert np.allclose(p.sum(0), 1.0, atol=1e-5) and p.min() >= 0 and p.max() <= 1
    assert np.array_equal(seg, (p[1] > p[0]).astype(np.uint8))
    # flipping the volume along x flips the result (tile grid is symmetric: steps [0,54] on 182)
    seg_f, p_f = full.predict_raw_volume_host(np.ascontiguousarray(raw[::-1]), do_mirroring=True)
    seg_t, p_t = full.predict_raw_volume_host(raw, do_mirroring=True)
    # (the Gaussian map is symmetric about index 64 of 0..127, not about 63.5, so this is approximate)
    assert np.abs(p_f[:, ::-1] - p_t).max() < 5e-2
    assert np.mean(seg_f[::-1] == seg_t) > 0.99


# ------------------------------------------------------------------------------------------------
# round 2: locally computed weight map, device inputs, resident ensemble, masked host call
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,gauss", [((40, 50, 45), True), ((33, 32, 47), True), ((40, 34, 32), False), ((32, 32, 32), True)])
def test_weight_map_is_bit_identical_to_the_tiles_own_sum(small, shape, gauss):
    tr = small[0]
    vol = torch.zeros(shape, device="cuda")
    agg = torch.zeros((2,) + shape, device="cuda"); wgt = torch.zeros(shape, device="cuda")
    tr.network.accumulate_tiles(vol, agg, wgt, 0.5, False, (), gauss)
    wm = tr.network.weight_map(shape, 0.5, gauss)
    assert torch.equal(wm, wgt)


def test_predict_3D_accepts_cuda_tensors(small):
    tr, net, plans = small
    for shape in [(40, 50, 45), (20, 32, 40)]:             # the second one is padded on axis 0
        data = O.synthetic_flair(shape, seed=11)
        data[0] = O.zscore_nnunet(data[0], np.where(data[0] != 0, 0, -1), True)
        seg_h, p_h = tr.predict_preprocessed_data_return_seg_and_softmax(data)
        seg_d, p_d = tr.predict_preprocessed_data_return_seg_and_softmax(torch.from_numpy(data).cuda())
        assert np.array_equal(p_h, p_d) and np.array_equal(seg_h, seg_d)


def test_resident_ensemble_equals_mean_of_separate_models():
    """config 5 machinery: k models resident with shared workspaces (dwmh_create_like), mean + argmax on the device."""
    from deepwmh_b200.parallel import make_ensemble, predict_volume_ensemble
    plans = small_plans()
    data = O.synthetic_flair((40, 34, 32), seed=12)
    data[0] = O.zscore_nnunet(data[0], np.where(data[0] != 0, 0, -1), True)
    nets = [O.build_benchmark_network(k, plans) for k in range(3)]
    trs = make_ensemble(plans, [n.state_dict() for n in nets], device=0, max_batch=8)
    seg, mean = predict_volume_ensemble(trs, torch.from_numpy(data[0]).cuda())
    singles = []
    for n in nets:
        tr, _ = _trainer(plans)
        tr.load_checkpoint_ram({"state_dict": n.state_dict()}, False)
        singles.append(tr.predict_preprocessed_data_return_seg_and_softmax(data)[1])
        tr.network.close()
    acc = np.zeros_like(singles[0])
    for s in singles:
        acc = acc + s * np.float32(1.0 / 3)                  # the device's fp32 axpy order (fma vs mul+add: <= 1 ulp)
    assert np.abs(mean.cpu().numpy() - acc).max() < 1e-6
    ref = np.mean(np.stack([O.OracleTrainer(plans, n).predict_preprocessed_data_return_seg_and_softmax(data)[1] for n in nets]), 0)
    _gate(ref.argmax(0), ref, seg.cpu().numpy(), mean.cpu().numpy(), tol=1e-2, agree=0.998, dice=0.995)
    # interleaving the models must not disturb each other's results (they share activation buffers, not weights)
    seg2, mean2 = predict_volume_ensemble(trs, torch.from_numpy(data[0]).cuda())
    assert torch.equal(mean, mean2) and torch.equal(seg, seg2)
    for tr in reversed(trs):
        tr.network.close()


def test_masked_host_call_uses_the_crop_mask(small):
    tr = small[0]
    raw = O.synthetic_flair((40, 50, 45), seed=13)[0]
    raw[10:14, 20:24, 20:24] = 0                                      # a hole inside the head: part of nnU-Net's filled mask
    from scipy.ndimage import binary_fill_holes
    mask = binary_fill_holes(raw != 0)
    segmask = np.where(mask, 0, -1).astype(np.int8)
    seg_m, p_m = tr.predict_raw_volume_host(raw, seg_mask=segmask)
    data = O.zscore_nnunet(raw, segmask, True)[None]
    seg_r, p_r = tr.predict_preprocessed_data_return_seg_and_softmax(data)
    # device z-score (fp64 statistics) vs numpy's fp32 one: inputs differ by ~1e-5, the fp16 pipeline amplifies a little
    assert np.abs(p_m - p_r).max() < 5e-3 and np.mean(seg_m == seg_r) > 0.998
    seg_2, p_2 = tr.predict_raw_volume_host(raw)                      # mask_mode 2 (vol != 0) differs inside the hole
    assert np.abs(p_m - p_2).max() > 1e-2
