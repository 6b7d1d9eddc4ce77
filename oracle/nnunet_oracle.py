"""Restatement of the nnU-Net v1 inference arithmetic that DeepWMH_predict runs.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): PARITY UNPINNED.

Every function names the upstream nnU-Net v1 symbol it restates ([U:...], the
un-vendored ``nnunet`` fork) and the line of
``/root/reference`` that forces it onto the hot path.  Built from stock
``torch.nn`` / numpy / scipy only; fp32 throughout (the reference's CPU path:
``torch.cuda.amp.autocast`` is a no-op on CPU).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from scipy.ndimage import gaussian_filter

__all__ = [
    "ConvDropoutNormNonlin", "StackedConvLayers", "Generic_UNet", "InitWeights_He",
    "benchmark_plans", "build_network", "build_benchmark_network", "count_parameters",
    "compute_steps_for_sliding_window", "get_gaussian", "pad_nd_image",
    "mirror_and_predict", "iter_predict_3D_tiled", "predict_3D_tiled", "predict_3D", "OracleTrainer",
    "zscore_nnunet", "zscore_deepwmh", "synthetic_flair", "hard_dice_binary",
    "parity_report", "forward_flops", "MIRROR_DIMS",
]


# --------------------------------------------------------------------------------------
# Network  [U:network_architecture/generic_UNet.py]; forced by
# `-tr nnUNetTrainerV2 -m 3d_fullres -p nnUNetPlansv2.1` (deepwmh/main/predict.py:134-136,153-155)
# --------------------------------------------------------------------------------------
class ConvDropoutNormNonlin(nn.Module):
    """conv(bias) -> InstanceNorm3d(eps 1e-5, affine, instance stats) -> LeakyReLU(0.01).
    Dropout p=0 in nnUNetTrainerV2 => no module.  Sub-module names kept (.conv/.instnorm/.lrelu)
    so genuine DeepWMH `model_best.model` state dicts load."""

    def __init__(self, cin: int, cout: int, kernel: Sequence[int], stride: Sequence[int]):
        super().__init__()
        pad = [1 if k == 3 else 0 for k in kernel]
        self.conv = nn.Conv3d(cin, cout, tuple(kernel), tuple(stride), tuple(pad), 1, bias=True)
        self.dropout = None
        self.instnorm = nn.InstanceNorm3d(cout, eps=1e-5, affine=True, momentum=0.1)
        self.lrelu = nn.LeakyReLU(negative_slope=1e-2, inplace=True)

    def forward(self, x):
        return self.lrelu(self.instnorm(self.conv(x)))


class StackedConvLayers(nn.Module):
    def __init__(self, cin, cout, num_convs, kernel, first_stride=None):
        super().__init__()
        one = [1] * len(kernel)
        blocks = [ConvDropoutNormNonlin(cin, cout, kernel, first_stride if first_stride is not None else one)]
        blocks += [ConvDropoutNormNonlin(cout, cout, kernel, one) for _ in range(num_convs - 1)]
        self.blocks = nn.Sequential(*blocks)
        self.output_channels = cout

    def forward(self, x):
        return self.blocks(x)


class Generic_UNet(nn.Module):
    """nnU-Net v1 `Generic_UNet` in the configuration nnUNetTrainerV2.initialize_network uses:
    convolutional pooling (strided first conv per stage), convolutional upsampling
    (ConvTranspose3d kernel=stride=pool, bias=False), 2 convs per stage, features
    min(32*2^d, 320), deep-supervision heads (1x1x1, bias=False), final_nonlin = identity,
    inference_apply_nonlin = softmax(dim=1)."""

    MAX_NUM_FILTERS_3D = 320

    def __init__(self, input_channels: int, base_num_features: int, num_classes: int,
                 pool_op_kernel_sizes: Sequence[Sequence[int]],
                 conv_kernel_sizes: Sequence[Sequence[int]], num_conv_per_stage: int = 2):
        super().__init__()
        num_pool = len(pool_op_kernel_sizes)
        assert len(conv_kernel_sizes) == num_pool + 1
        self.num_classes = num_classes
        self.do_ds = True
        self._deep_supervision = True
        self.inference_apply_nonlin = lambda x: F.softmax(x, 1)
        self.pool_op_kernel_sizes = [list(p) for p in pool_op_kernel_sizes]
        self.conv_kernel_sizes = [list(k) for k in conv_kernel_sizes]
        self._gaussian_3d = None
        self._patch_size_for_gaussian_3d = None

        ctx, loc, tu, heads = [], [], [], []
        fin, fout = input_channels, base_num_features
        for d in range(num_pool):
            stride = self.pool_op_kernel_sizes[d - 1] if d > 0 else None
            ctx.append(StackedConvLayers(fin, fout, num_conv_per_stage, self.conv_kernel_sizes[d], stride))
            fin = fout
            fout = min(int(np.round(fout * 2)), self.MAX_NUM_FILTERS_3D)
        # bottleneck: two single-conv stacks, first one strided
        ctx.append(nn.Sequential(
            StackedConvLayers(fin, fout, num_conv_per_stage - 1, self.conv_kernel_sizes[num_pool],
                              self.pool_op_kernel_sizes[-1]),
            StackedConvLayers(fout, fout, 1, self.conv_kernel_sizes[num_pool])))
        cur = fout
        for u in range(num_pool):
            skip = ctx[-(2 + u)].output_channels
            k = self.pool_op_kernel_sizes[-(u + 1)]
            tu.append(nn.ConvTranspose3d(cur, skip, tuple(k), tuple(k), bias=False))
            ck = self.conv_kernel_sizes[-(u + 1)]
            loc.append(nn.Sequential(StackedConvLayers(2 * skip, skip, num_conv_per_stage - 1, ck),
                                     StackedConvLayers(skip, skip, 1, ck)))
            heads.append(nn.Conv3d(skip, num_classes, 1, 1, 0, 1, 1, bias=False))
            cur = skip
        self.conv_blocks_context = nn.ModuleList(ctx)
        self.conv_blocks_localization = nn.ModuleList(loc)
        self.tu = nn.ModuleList(tu)
        self.seg_outputs = nn.ModuleList(heads)
        self.apply(InitWeights_He(1e-2))

    def forward(self, x):
        skips, outs = [], []
        for d in range(len(self.conv_blocks_context) - 1):
            x = self.conv_blocks_context[d](x)
            skips.append(x)
        x = self.conv_blocks_context[-1](x)
        for u in range(len(self.tu)):
            x = self.tu[u](x)
            x = torch.cat((x, skips[-(u + 1)]), dim=1)          # upsampled first, skip second
            x = self.conv_blocks_localization[u](x)
            outs.append(self.seg_outputs[u](x))                 # final_nonlin = identity
        if self._deep_supervision and self.do_ds:
            return tuple([outs[-1]] + outs[:-1][::-1])
        return outs[-1]


class InitWeights_He:
    """[U:network_architecture/initialization.py] kaiming_normal_(a=neg_slope) on conv and
    transposed-conv weights, bias <- 0."""

    def __init__(self, neg_slope: float = 1e-2):
        self.neg_slope = neg_slope

    def __call__(self, m):
        if isinstance(m, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            m.weight = nn.init.kaiming_normal_(m.weight, a=self.neg_slope)
            if m.bias is not None:
                m.bias = nn.init.constant_(m.bias, 0)


def benchmark_plans(patch_size=(128, 128, 128), num_pool=5, base_num_features=32) -> Dict:
    """The canonical benchmark instance of SURVEY.md section 8d (a `plans`-shaped dict holding
    exactly the keys `process_plans` reads for stage 0 of 3d_fullres)."""
    return {
        "num_modalities": 1, "num_classes": 1,            # plans['num_classes'] excludes background
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


def build_network(plans: Dict) -> Generic_UNet:
    st = plans["plans_per_stage"][max(plans["plans_per_stage"].keys())]
    return Generic_UNet(plans["num_modalities"], plans["base_num_features"], plans["num_classes"] + 1,
                        st["pool_op_kernel_sizes"], st["conv_kernel_sizes"], 2)


def build_benchmark_network(model_index: int = 0, plans: Optional[Dict] = None,
                            randomize_affine: bool = True) -> Generic_UNet:
    """Random-init weights of BASELINE.json: torch.manual_seed(1234+k) -> InitWeights_He(1e-2);
    IN affine and conv biases additionally randomised (gamma~U(0.5,1.5), beta~N(0,0.1),
    bias~N(0,0.1)) so those code paths carry signal."""
    torch.manual_seed(1234 + model_index)
    net = build_network(plans if plans is not None else benchmark_plans())
    if randomize_affine:
        g = torch.Generator().manual_seed(4321 + model_index)
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, nn.InstanceNorm3d):
                    m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                elif isinstance(m, nn.Conv3d) and m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    net.eval()
    net.do_ds = False
    return net


def count_parameters(net: nn.Module) -> int:
    return sum(p.numel() for p in net.parameters())


def forward_flops(plans: Dict) -> float:
    """2*MAC over conv + transposed conv + final 1x1x1 head for one patch forward (the
    algorithmic work BASELINE.md section 3 quotes: 954.46 GFLOP at 128^3)."""
    st = plans["plans_per_stage"][max(plans["plans_per_stage"].keys())]
    pools, kers = st["pool_op_kernel_sizes"], st["conv_kernel_sizes"]
    shape = np.array(st["patch_size"], dtype=np.int64)
    feats = [plans["base_num_features"]]
    for _ in pools:
        feats.append(min(int(np.round(feats[-1] * 2)), Generic_UNet.MAX_NUM_FILTERS_3D))
    mac, cin, shapes = 0, plans["num_modalities"], []
    for d in range(len(pools) + 1):
        if d > 0:
            shape = shape // np.array(pools[d - 1])
        k = int(np.prod(kers[d]))
        mac += int(np.prod(shape)) * k * (cin * feats[d] + feats[d] * feats[d])
        cin = feats[d]
        shapes.append(shape.copy())
    for u in range(len(pools)):
        skip = feats[-(2 + u)]
        shape = shapes[-(2 + u)]
        k = int(np.prod(kers[-(u + 1)]))
        mac += int(np.prod(shape)) * cin * skip                        # k=s transposed conv
        mac += int(np.prod(shape)) * k * (2 * skip * skip + skip * skip)
        cin = skip
    mac += int(np.prod(shapes[0])) * cin * (plans["num_classes"] + 1)  # last head only
    return 2.0 * mac


# --------------------------------------------------------------------------------------
# Sliding-window helpers  [U:network_architecture/neural_network.py]
# --------------------------------------------------------------------------------------
def compute_steps_for_sliding_window(patch_size, image_size, step_size: float) -> List[List[int]]:
    """[U::_compute_steps_for_sliding_window]"""
    assert all(i >= p for i, p in zip(image_size, patch_size)), "image must be >= patch"
    assert 0 < step_size <= 1
    steps = []
    for img, p in zip(image_size, patch_size):
        target = p * step_size
        n = int(np.ceil((img - p) / target)) + 1
        max_step = img - p
        actual = max_step / (n - 1) if n > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(n)])
    return steps


def get_gaussian(patch_size, sigma_scale: float = 1.0 / 8) -> np.ndarray:
    """[U::_get_gaussian] unit impulse at patch//2 -> scipy gaussian_filter(sigma=patch*scale,
    mode constant) -> /max -> fp32 -> zeros replaced by the smallest non-zero."""
    tmp = np.zeros(patch_size)
    tmp[tuple(i // 2 for i in patch_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in patch_size], 0, mode="constant", cval=0)
    g = (g / np.max(g) * 1).astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def pad_nd_image(image: np.ndarray, new_shape, mode="constant", kwargs=None):
    """[U:batchgenerators.augmentations.utils::pad_nd_image] symmetric pad of the trailing
    len(new_shape) axes up to new_shape (below = d//2, above = d//2 + d%2); returns
    (padded, slicer-that-undoes-it)."""
    kwargs = kwargs or {"constant_values": 0}
    nd = len(new_shape)
    old = np.array(image.shape[-nd:])
    tgt = np.maximum(np.array(new_shape), old)
    diff = tgt - old
    below, above = diff // 2, diff // 2 + diff % 2
    pad = [[0, 0]] * (image.ndim - nd) + [[int(b), int(a)] for b, a in zip(below, above)]
    res = np.pad(image, pad, mode, **kwargs) if diff.any() else image
    slicer = tuple(slice(p[0], res.shape[i] - p[1]) for i, p in enumerate(pad))
    return res, slicer


# flip dims (on the [1,C,x,y,z] tensor) per mirror index m; [U::_internal_maybe_mirror_and_pred_3D]
MIRROR_DIMS = {0: (), 1: (4,), 2: (3,), 3: (4, 3), 4: (2,), 5: (4, 2), 6: (3, 2), 7: (4, 3, 2)}
_MIRROR_NEEDS = {0: (), 1: (2,), 2: (1,), 3: (2, 1), 4: (0,), 5: (0, 2), 6: (0, 1), 7: (0, 1, 2)}


@torch.no_grad()
def mirror_and_predict(net: Generic_UNet, x: torch.Tensor, mirror_axes=(0, 1, 2),
                       do_mirroring=True, mult: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[U::_internal_maybe_mirror_and_pred_3D]  x [1,C,px,py,pz] -> [1,classes,px,py,pz] fp32:
    sum_m (1/2^len(axes)) * unflip(softmax(net(flip_m x))), then *= gaussian."""
    # the reference moves the tile to the network's device and brings the result back (all_in_gpu=False); on a CPU
    # network both moves are no-ops.  A CUDA network is only used by the GPU parity tests (fp32, TF32 off).
    dev = next(net.parameters()).device if hasattr(net, "parameters") else x.device
    x = x.to(dev)
    if mult is not None:
        mult = mult.to(dev)
    result = torch.zeros([1, net.num_classes] + list(x.shape[2:]), dtype=torch.float32, device=dev)
    n_mirror = 8 if do_mirroring else 1
    num_results = 2 ** len(mirror_axes) if do_mirroring else 1
    for m in range(n_mirror):
        if not all(a in mirror_axes for a in _MIRROR_NEEDS[m]):
            continue
        dims = MIRROR_DIMS[m]
        xin = torch.flip(x, dims) if dims else x
        pred = net.inference_apply_nonlin(net(xin))
        if dims:
            pred = torch.flip(pred, dims)
        result += 1 / num_results * pred
    if mult is not None:
        result[:, :] *= mult
    return result.cpu()


@torch.no_grad()
def iter_predict_3D_tiled(net: Generic_UNet, x: np.ndarray, step_size: float, do_mirroring: bool,
                          mirror_axes, patch_size, use_gaussian: bool, pad_border_mode="constant",
                          pad_kwargs=None, return_buffers=False):
    """[U::_internal_predict_3D_3Dconv_tiled] default branch (all_in_gpu=False: host fp32 numpy aggregation) as a
    generator: yields ("tile", index, n_tiles) after every tile and finally ("done", result).  The generator form lets
    bench.py's CPU arm time a bounded number of tiles of the very same code path."""
    assert x.ndim == 4, "x must be (c, x, y, z)"
    data, slicer = pad_nd_image(x, patch_size, pad_border_mode, pad_kwargs)
    shape = data.shape
    steps = compute_steps_for_sliding_window(patch_size, shape[1:], step_size)
    num_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    if use_gaussian and num_tiles > 1:
        if net._gaussian_3d is None or tuple(net._patch_size_for_gaussian_3d) != tuple(patch_size):
            net._gaussian_3d = get_gaussian(patch_size, 1.0 / 8)
            net._patch_size_for_gaussian_3d = tuple(patch_size)
        add_nb = net._gaussian_3d
        mult = torch.from_numpy(net._gaussian_3d)
    else:
        add_nb = np.ones(patch_size, dtype=np.float32)
        mult = None
    agg = np.zeros([net.num_classes] + list(shape[1:]), dtype=np.float32)
    nb = np.zeros([net.num_classes] + list(shape[1:]), dtype=np.float32)
    done = 0
    for lx in steps[0]:
        for ly in steps[1]:
            for lz in steps[2]:
                sl = (slice(lx, lx + patch_size[0]), slice(ly, ly + patch_size[1]), slice(lz, lz + patch_size[2]))
                tile = torch.from_numpy(np.ascontiguousarray(data[(None, slice(None)) + sl]))
                pred = mirror_and_predict(net, tile, mirror_axes, do_mirroring, mult)[0].numpy()
                agg[(slice(None),) + sl] += pred
                nb[(slice(None),) + sl] += add_nb
                done += 1
                yield ("tile", done, num_tiles)
    sl_out = (slice(0, agg.shape[0]),) + tuple(slicer[1:])
    agg, nb = agg[sl_out], nb[sl_out]
    if return_buffers:
        yield ("done", (agg.copy(), nb.copy()))
        return
    probs = agg / nb
    yield ("done", (probs.argmax(0), probs))


def predict_3D_tiled(net: Generic_UNet, x: np.ndarray, step_size: float, do_mirroring: bool,
                     mirror_axes, patch_size, use_gaussian: bool, pad_border_mode="constant",
                     pad_kwargs=None, return_buffers=False):
    """[U::_internal_predict_3D_3Dconv_tiled]  x [C,X,Y,Z] -> (seg int64 [X,Y,Z], softmax fp32 [classes,X,Y,Z])."""
    out = None
    for kind, *payload in iter_predict_3D_tiled(net, x, step_size, do_mirroring, mirror_axes, patch_size, use_gaussian,
                                                pad_border_mode, pad_kwargs, return_buffers):
        if kind == "done":
            out = payload[0]
    return out


def predict_3D(net: Generic_UNet, x: np.ndarray, do_mirroring: bool, mirror_axes=(0, 1, 2),
               use_sliding_window=False, step_size=0.5, patch_size=None, regions_class_order=None,
               use_gaussian=False, pad_border_mode="constant", pad_kwargs=None, all_in_gpu=False,
               verbose=True, mixed_precision=True):
    """[U::SegmentationNetwork.predict_3D] guards + dispatch (only the tiled 3-D conv branch is
    on DeepWMH's path: predict.py:153-155 passes no flag that would select another)."""
    assert step_size <= 1, "step_size must be smaller than 1"
    assert x.ndim == 4, "data must have shape (c,x,y,z)"
    if pad_kwargs is None:
        pad_kwargs = {"constant_values": 0}
    if len(mirror_axes):
        assert max(mirror_axes) <= 2, "mirror axes"
    assert use_sliding_window and regions_class_order is None and not all_in_gpu
    was_training = net.training
    net.eval()
    out = predict_3D_tiled(net, x, step_size, do_mirroring, mirror_axes, patch_size, use_gaussian,
                           pad_border_mode, pad_kwargs)
    net.train(was_training)
    return out


class OracleTrainer:
    """[U:training/network_training/nnUNetTrainerV2.py] the slice of the trainer surface the
    predictor uses: predict_preprocessed_data_return_seg_and_softmax / load_checkpoint_ram."""

    def __init__(self, plans: Optional[Dict] = None, network: Optional[Generic_UNet] = None):
        self.plans = plans if plans is not None else benchmark_plans()
        st = self.plans["plans_per_stage"][max(self.plans["plans_per_stage"].keys())]
        self.patch_size = np.array(st["patch_size"]).astype(int)
        self.num_classes = self.plans["num_classes"] + 1
        self.data_aug_params = {"do_mirror": True, "mirror_axes": (0, 1, 2)}
        self.network = network if network is not None else build_network(self.plans)
        self.network.do_ds = False

    def load_checkpoint_ram(self, checkpoint: Dict, train: bool = False):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in checkpoint["state_dict"].items()}
        self.network.load_state_dict(sd)

    def predict_preprocessed_data_return_seg_and_softmax(
            self, data: np.ndarray, do_mirroring: bool = True, mirror_axes=None,
            use_sliding_window: bool = True, step_size: float = 0.5, use_gaussian: bool = True,
            pad_border_mode: str = "constant", pad_kwargs: dict = None, all_in_gpu: bool = False,
            verbose: bool = True, mixed_precision=True):
        if pad_border_mode == "constant" and pad_kwargs is None:
            pad_kwargs = {"constant_values": 0}
        if do_mirroring and mirror_axes is None:
            mirror_axes = self.data_aug_params["mirror_axes"]
        if do_mirroring:
            assert self.data_aug_params["do_mirror"], "cannot mirror at test time if not trained with it"
        ds, self.network.do_ds = self.network.do_ds, False
        self.network.eval()
        ret = predict_3D(self.network, data, do_mirroring, mirror_axes or (), use_sliding_window, step_size,
                         tuple(int(i) for i in self.patch_size), None, use_gaussian, pad_border_mode,
                         pad_kwargs, all_in_gpu, verbose, mixed_precision)
        self.network.do_ds = ds
        return ret


# --------------------------------------------------------------------------------------
# Intensity normalisation
# --------------------------------------------------------------------------------------
def zscore_nnunet(data: np.ndarray, seg: Optional[np.ndarray], use_mask_for_norm: bool) -> np.ndarray:
    """[U:preprocessing/preprocessing.py::GenericPreprocessor.resample_and_normalize], non-CT
    scheme (forced by modality name 'Modality_01', deepwmh/pipeline/DCNN_multistage.py:39-41,82).
    data [X,Y,Z] fp32 (one channel); seg>=0 marks the nonzero-crop mask.  Population std."""
    out = data.astype(np.float32, copy=True)
    if use_mask_for_norm:
        mask = seg >= 0
        mn = out[mask].mean()
        sd = out[mask].std()
        out[mask] = (out[mask] - mn) / (sd + 1e-8)
        out[mask == 0] = 0
    else:
        mn = out.mean()
        sd = out.std()
        out = (out - mn) / (sd + 1e-8)
    return out


def zscore_deepwmh(data: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
    """In-tree sibling `z_score` (deepwmh/analysis/image_ops.py:172-179 with masked_mean/std
    :13-21): statistics over mask>0.5, std floored at 1e-5, applied to ALL voxels."""
    if mask is None:
        mn, sd = np.mean(data), np.std(data)
    else:
        m = mask > 0.5
        mn, sd = data[m].mean(), data[m].std()
    sd = max(float(sd), 0.00001)
    return (data - mn) / sd


# --------------------------------------------------------------------------------------
# Synthetic inputs and parity metrics
# --------------------------------------------------------------------------------------
def synthetic_flair(shape=(182, 218, 182), seed: int = 0) -> np.ndarray:
    """fp32 [1,X,Y,Z]: ellipsoid head (semi-axes 0.45*shape) of clip(N(100,25),1,.) tissue plus ~40
    Gaussian hyper-intense blobs; exactly 0 outside the head."""
    rng = np.random.default_rng(seed)
    gx, gy, gz = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    c = [(s - 1) / 2.0 for s in shape]
    r = [0.45 * s for s in shape]
    head = ((gx - c[0]) / r[0]) ** 2 + ((gy - c[1]) / r[1]) ** 2 + ((gz - c[2]) / r[2]) ** 2 <= 1.0
    vol = np.clip(rng.normal(100.0, 25.0, size=shape), 1.0, None).astype(np.float32)
    for _ in range(40):
        ctr = [rng.uniform(0.25 * s, 0.75 * s) for s in shape]
        sig = rng.uniform(1.0, 4.0)
        amp = rng.uniform(60.0, 150.0)
        rad = int(math.ceil(4 * sig))
        lo = [max(0, int(ctr[i]) - rad) for i in range(3)]
        hi = [min(shape[i], int(ctr[i]) + rad + 1) for i in range(3)]
        sub = (slice(lo[0], hi[0]), slice(lo[1], hi[1]), slice(lo[2], hi[2]))
        d2 = (gx[sub] - ctr[0]) ** 2 + (gy[sub] - ctr[1]) ** 2 + (gz[sub] - ctr[2]) ** 2
        vol[sub] += (amp * np.exp(-d2 / (2 * sig * sig))).astype(np.float32)
    vol[~head] = 0.0
    return vol[None].astype(np.float32)


def hard_dice_binary(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """Dice definition of deepwmh/analysis/metrics.py:26-32."""
    a = (y_true > 0.5).astype(np.float32)
    b = (y_pred > 0.5).astype(np.float32)
    return float(2 * np.sum(a * b) / (np.sum(a) + np.sum(b) + 0.000001))


def parity_report(seg_ref, prob_ref, seg_new, prob_new) -> Dict[str, float]:
    """The three numbers of BASELINE.json's parity gate."""
    return {
        "softmax_max_abs": float(np.max(np.abs(prob_ref.astype(np.float64) - prob_new.astype(np.float64)))),
        "argmax_agree": float(np.mean(seg_ref == seg_new)),
        "dice": hard_dice_binary(seg_ref, seg_new),
        "fg_frac_ref": float(np.mean(seg_ref > 0)),
    }
