"""CPU restatement of the arithmetic that lives IN /root/reference either side of the network.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py for who may import this).

PARITY PINNED: unlike the nnU-Net engine, these functions are in the reference tree and run in the authoring
container; tests/golden/make_golden_intree.py executes the reference's own code on seeded inputs and commits the
outputs as tests/golden/intree_v1.npz; tests/test_intree_oracle.py checks every function below against them.

All functions follow the reference's float64 arithmetic (numpy promotes the float32 volumes against the float64
statistics); inputs are numpy arrays [X, Y, Z].
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence

import numpy as np
from scipy.ndimage import binary_erosion, label, median_filter, zoom

SQRT_2PI_REF = 2.506   # the reference's rounded sqrt(2 pi), deepwmh/analysis/lesion_analysis.py:103


def masked_mean_std(data, mask):
    """masked_mean / masked_std, deepwmh/analysis/image_ops.py:13-21 (population std over mask > 0.5)."""
    m = np.asarray(mask) > 0.5
    v = np.asarray(data)[m].astype(np.float64)
    return float(v.mean()), float(v.std())


def z_score(data, mask=None):
    """z_score, deepwmh/analysis/image_ops.py:172-179: statistics over the mask, std floored at 1e-5, ALL voxels
    normalised."""
    data = np.asarray(data)
    if mask is None:
        mn, sd = float(np.mean(data, dtype=np.float64)), float(np.std(data, dtype=np.float64))
    else:
        mn, sd = masked_mean_std(data, mask)
    return (data.astype(np.float64) - mn) / max(sd, 0.00001)


def tissue_min_fill(x, brain):
    """lesion_analysis.py:150-151 / 160-161: voxels outside the rough brain mask <- minimum inside it."""
    b = np.asarray(brain) >= 0.5
    return np.where(b, x, np.asarray(x)[b].min())


def remove_sparks(mask, min_volume=3):
    """remove_sparks, image_ops.py:325-344: 6-connected components with fewer than min_volume voxels are dropped."""
    lab, n = label((np.asarray(mask) > 0.5).astype("int"))
    sizes = np.bincount(lab.ravel(), minlength=n + 1)
    keep = sizes >= min_volume
    keep[0] = False
    return keep[lab].astype("int")


def spark_min_volume(voxel_size: Sequence[float]) -> int:
    """The voxel-size rule of remove_3mm_sparks, image_ops.py:346-367."""
    vs = [float(v) for v in voxel_size]
    if max(vs) / min(vs) > 3.0:
        return 3
    return max(2, int(np.around(3.0 / (vs[0] * vs[1] * vs[2]))))


def remove_3mm_sparks(mask, voxel_size):
    return remove_sparks(mask, spark_min_volume(voxel_size))


def softmax_masking(x, m):
    """_parallel_softmax_masking, deepwmh/pipeline/DCNN_multistage.py:102-109 (float32 in, float32 out)."""
    return 1 - (np.asarray(m, np.float32) * (1 - np.asarray(x, np.float32)))


def ensembling(masked: Sequence[np.ndarray], voxel_size):
    """_parallel_ensembling, DCNN_multistage.py:111-125 -> (field float32, refined label)."""
    field = np.zeros(masked[0].shape).astype("float32")
    for y in masked:
        field += np.asarray(y).astype("float32")
    field = field / len(masked)
    return field, remove_3mm_sparks((field < 0.5).astype("float32"), list(voxel_size))


def hard_dice_binary(y_true, y_pred):
    """deepwmh/analysis/metrics.py:26-32."""
    a = (np.asarray(y_true) > 0.5).astype("float32")
    b = (np.asarray(y_pred) > 0.5).astype("float32")
    return 2 * np.sum(a * b) / (np.sum(a) + np.sum(b) + 0.000001)


def _stack_masked(xs, masks):
    rows = [np.asarray(x) for x in xs]
    if masks is not None:
        rows = [np.where(np.asarray(m) < 0.5, np.nan, x) for x, m in zip(rows, masks)]
    return np.stack(rows)


def group_mean(xs: Sequence[np.ndarray], masks=None):
    """group_mean, image_ops.py:216-231: voxelwise mean over the reference images (in the dtype of the inputs, as
    np.nanmean of the stacked rows does: float32 volumes are averaged in float32); voxels with mask < 0.5 are left out
    (NaN where no image remains)."""
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.nanmean(_stack_masked(xs, masks), axis=0)


def group_std(xs: Sequence[np.ndarray], masks=None):
    """group_std, image_ops.py:199-214: voxelwise population std (dtype of the inputs), masks as above."""
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.nanstd(_stack_masked(xs, masks), axis=0)


def nll(x_prime, x_refs: List[np.ndarray], min_std: Optional[float] = None, side: Optional[str] = None,
        return_all: bool = False, use_mask: bool = False):
    """nll, deepwmh/analysis/lesion_analysis.py:84-113; use_mask: every reference contributes only where it exceeds its
    own Otsu threshold (:87-92; threshold_otsu below is the unpinned restatement of skimage's)."""
    assert side in (None, "+", "-")
    masks = [np.where(x > threshold_otsu(x), 1, 0) for x in x_refs] if use_mask else None
    mu, sigma = group_mean(x_refs, masks), group_std(x_refs, masks)
    sigma = sigma + 1e-6 if min_std is None else np.where(sigma < min_std, min_std, sigma)
    x = np.asarray(x_prime)
    with np.errstate(all="ignore"):
        an = np.power(x - mu, 2) / (2 * np.power(sigma, 2)) + np.log(sigma * SQRT_2PI_REF)
    an = np.nan_to_num(an, nan=0.0)
    if side == "+":
        an = an * (x > mu)
    elif side == "-":
        an = an * (x < mu)
    return (an, mu, sigma) if return_all else an


def mean_std_grid(data, patch_size, mask=None):
    """mean_std_grid (order 1), image_ops.py:56-170: masked mean / std of half-overlapping blocks on the zero-padded
    volume, zero-bordered grid, linear zoom by the step, cut back at offset step // 2."""
    data = np.asarray(data, np.float64)
    ps = [int(2 * np.ceil(p / 2)) for p in patch_size]
    st = [p // 2 for p in ps]
    shp = data.shape
    pshape = [ps[a] * int(np.ceil(shp[a] / ps[a])) for a in range(3)]
    pd = np.zeros(pshape)
    pd[:shp[0], :shp[1], :shp[2]] = data
    pm = None
    if mask is not None:
        pm = np.zeros(pshape, bool)
        pm[:shp[0], :shp[1], :shp[2]] = np.asarray(mask) > 0.5
    g = [pshape[a] // st[a] for a in range(3)]
    mg, sg = np.zeros([v + 2 for v in g]), np.zeros([v + 2 for v in g])
    for i in range(g[0]):
        for j in range(g[1]):
            for k in range(g[2]):
                sl = tuple(slice(c * st[a], c * st[a] + ps[a]) for a, c in enumerate((i, j, k)))   # clipped at the end
                blk = pd[sl]
                if pm is not None:
                    sel = blk[pm[sl]]
                    mu, sd = (sel.mean(), sel.std()) if sel.size else (0.0, 0.00001)
                else:
                    mu, sd = blk.mean(), max(blk.std(), 0.00001)
                mg[i + 1, j + 1, k + 1], sg[i + 1, j + 1, k + 1] = mu, sd
    out = []
    for grid in (mg, sg):
        z = zoom(grid, st, order=1)
        o = [s // 2 for s in st]
        z = z[o[0]:o[0] + g[0] * st[0], o[1]:o[1] + g[1] * st[1], o[2]:o[2] + g[2] * st[2]]
        out.append(z[:shp[0], :shp[1], :shp[2]])
    return out[0], out[1]


def median_kernel(voxel_size: Sequence[float]) -> List[int]:
    """Kernel-size rule of median_3mm, image_ops.py:378-421: int(3 mm / voxel) per axis, at least 3; for thick slices
    (max / min > 4) the filter is 2-D, i.e. size 1 along the thick axis."""
    vs = [float(v) for v in voxel_size]
    ks = [max(3, int(3.0 / v)) for v in vs]
    if max(vs) / min(vs) > 4.0:
        ks[int(np.argmax(vs))] = 1
    return ks


def median_3mm(data, voxel_size):
    return median_filter(np.asarray(data), size=median_kernel(voxel_size), mode="constant", cval=0)


def component_filtering(mask, voxel_size):
    """component_filtering, image_ops.py:253-306: for every slice of every filtered orientation keep the largest
    4-connected component of the eroded slice (ties: the first label in raster order); thin-slice data (max / min voxel
    size <= 3) filters all three orientations, thick-slice data only the slices across the thick axis while the other two
    orientations contribute the mask itself; result = (sum of the three volumes) > 0.5."""
    mask = np.asarray(mask)
    vs = [float(v) for v in voxel_size]
    axes = [int(np.argmax(vs))] if max(vs) / min(vs) > 3 else [0, 1, 2]
    total = np.zeros(mask.shape, np.float32)
    for ax in range(3):
        if ax not in axes:
            total += mask.astype(np.float32)
            continue
        vol = np.zeros(mask.shape, np.float32)
        for s_ in range(mask.shape[ax]):
            sl = [slice(None)] * 3
            sl[ax] = s_
            lab, n = label(binary_erosion(mask[tuple(sl)]))
            if n:
                sizes = np.bincount(lab.ravel(), minlength=n + 1)[1:]
                vol[tuple(sl)] = lab == 1 + int(np.argmax(sizes))      # argmax: first of equal maxima
        total += vol
    return (total > 0.5).astype(np.float32)


def threshold_otsu(values, nbins=256):
    """skimage.filters.threshold_otsu (third-party, NOT vendored under /root/reference and absent here; scikit-image >= 0.16
    published algorithm, PARITY UNPINNED): 256 equal-width bins over [min, max] (numpy.histogram), threshold = centre of
    the bin that maximises the between-class variance w1 * w2 * (m1 - m2)^2; a constant image returns its value.
    Reference call sites: lesion_analysis.py:90,145-146, image_ops.py:308-323."""
    v = np.asarray(values, np.float64).reshape(-1)
    if np.all(v == v[0]):
        return float(v[0])
    counts, edges = np.histogram(v, bins=nbins)
    counts = counts.astype(np.float64)
    centers = (edges[:-1] + edges[1:]) / 2.0
    w1, w2 = np.cumsum(counts), np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        m1 = np.cumsum(counts * centers) / w1
        m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
    return float(centers[int(np.argmax(w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2))])


def valid_score_mask(x_prime_raw, brain, apply_otsu=True):
    """lesion_analysis.py:142-148."""
    x = z_score(x_prime_raw, brain)
    b = (np.asarray(brain) >= 0.5)
    if not apply_otsu:
        return x, b.astype(np.float32)
    thr = threshold_otsu(np.where(b, x, x.min()))
    return x, (b * (x > thr)).astype(np.float32)


def nll_anomaly_arrays(target, refs, brain, valid, patch, min_std=0.03, side="+", mean_correction=True):
    """The array part of nll_analysis, lesion_analysis.py:142-176 (after the masks are known)."""
    x_prime = tissue_min_fill(z_score(target, brain), brain)
    x_i = [tissue_min_fill(z_score(r, brain), brain) for r in refs]
    mu_p, _ = mean_std_grid(x_prime, patch, mask=valid)
    if mean_correction:
        x_i = [x - mean_std_grid(x, patch, mask=valid)[0] + mu_p for x in x_i]
    an, xm, xs = nll(x_prime, x_i, min_std=min_std, side=side, return_all=True)
    return {"x_prime": x_prime, "local_mu": mu_p, "anomaly": an * valid, "mean": xm, "std": xs, "refs": x_i}


def average_contiguous_labels(labels):
    """image_ops.py:23-38: voxelwise majority label over the maps (argmax of the label histogram: first maximum)."""
    n = 1 + max(int(np.max(l)) for l in labels)
    votes = np.stack([sum((np.asarray(l).astype("int") == c).astype(np.float32) for l in labels) for c in range(n)])
    return np.argmax(votes, axis=0)


def hist_curve(data, bins, log_y=False):
    """lesion_analysis.py:40-50."""
    hist, edges = np.histogram(data, bins=bins)
    centers = (edges[:-1] + edges[1:]) / 2
    if log_y:
        hist = np.log10(np.where(hist == 0, 0.001, hist))
        hist = np.where(hist < 0, 0, hist)
    return centers, hist


def nll_analysis_arrays(x, refs, label1, label2, voxel_size, intensity_prior=None, apply_otsu=False):
    """nll_analysis, lesion_analysis.py:115-281, on arrays -> (anomaly, m_valid, curve_x, curve_y, curve_r, threshold, extras)."""
    vox = [float(v) for v in voxel_size]
    patch = [int(np.ceil(50.0 / v)) for v in vox]
    rough = (np.mean(np.stack([(np.asarray(m) > 0.5).astype(np.float32) for m in label1]), axis=0) > 0.5).astype(np.float32)
    _, valid = valid_score_mask(x, rough, apply_otsu)
    r = nll_anomaly_arrays(x, refs, rough, valid, patch, min_std=0.03, side=intensity_prior)
    anomaly = r["anomaly"] * component_filtering(valid, vox)
    an_refs = [nll(s, r["refs"], min_std=0.03, side=intensity_prior) * valid for s in r["refs"]]
    means = [a[(valid > 0.5) & (a > 0)].mean() for a in an_refs]
    bins = np.linspace(0.0, 400 * (np.mean(means) / 4), 401)
    cx, cy = hist_curve(anomaly, bins, True)
    rs = [hist_curve(a, bins, True)[1] for a in an_refs]
    crossings = [cx[max(j for j in range(1, len(c)) if c[j] > 0.01)] for c in rs]
    thr = float(np.median(crossings))
    avg = average_contiguous_labels(label2)
    anomaly = anomaly * (avg > 0.5)
    anomaly = np.where((avg > 1.5) & (avg < 2.5), median_3mm(anomaly, vox), anomaly)
    tissue = sum((np.asarray(t) > 0.5).astype(np.float32) for t in label2) > len(label2) / 2
    anomaly = anomaly * tissue
    return anomaly, valid, cx, cy, np.mean(rs, axis=0), thr, {"averaged_label": avg, "rough_brain": rough, "mean": r["mean"],
                                                               "std": r["std"], "x_prime": r["x_prime"], "local_mu": r["local_mu"]}
