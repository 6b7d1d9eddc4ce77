"""Multi-GPU drivers for the path: one process per GPU, `torch.distributed`.

* cohort sharding  -- subjects are independent: subject i -> rank i mod world; no collective on
  the data path (only host-side gathers of names / timings).
* tile sharding    -- the tiles of ONE large volume are split into contiguous ranges per rank; every
  rank accumulates into its own full-size fp32 agg/wgt buffers, then a single all-reduce (NCCL over
  NVLink on GPUs; gloo in the CPU tests) sums them before finalize.
* ensemble         -- k models looped inside each rank per subject; softmax mean on device (a13).

Nothing in the reference corresponds to this file (it is single-GPU: deepwmh/main/predict.py:150).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_cohort(n_cases: int, rank: int, world: int) -> List[int]:
    """Static round-robin: equal-size volumes, so subject i goes to rank i mod world."""
    return list(range(rank, n_cases, world))


def shard_tiles(n_tiles: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) ranges in x-outer tile order (neighbouring tiles stay on one GPU, so the
    region each rank touches is a slab).  The first n_tiles % world ranks get one extra tile."""
    base, rem = divmod(n_tiles, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_buffers(agg: torch.Tensor, wgt: torch.Tensor, group=None):
    """The one exchange step of the tile-sharded mode: sum agg [2,X,Y,Z] and wgt [X,Y,Z] over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(wgt, op=dist.ReduceOp.SUM, group=group)


def predict_volume_tile_sharded(trainer, data: np.ndarray, do_mirroring=True, mirror_axes=(0, 1, 2),
                                step_size=0.5, use_gaussian=True, group=None):
    """Tile-sharded predict_preprocessed_data_return_seg_and_softmax: every rank passes the same
    (c,x,y,z) array and gets the same (seg, softmax) back as device tensors."""
    from .predictor import pad_nd_image
    net = trainer.network
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    padded, slicer = pad_nd_image(np.asarray(data, dtype=np.float32), net.patch_size)
    vol = torch.from_numpy(np.ascontiguousarray(padded[0])).to(net.device)
    X, Y, Z = vol.shape
    agg = torch.zeros((2, X, Y, Z), dtype=torch.float32, device=net.device)
    wgt = torch.zeros((X, Y, Z), dtype=torch.float32, device=net.device)
    n_tiles = net.num_tiles((X, Y, Z), step_size)
    b, e = shard_tiles(n_tiles, rank, world)
    # use_gaussian must be decided on the GLOBAL tile count (the reference: num_tiles > 1)
    if e > b:
        net.accumulate_tiles(vol, agg, wgt, step_size, do_mirroring, mirror_axes, use_gaussian and n_tiles > 1, b, e)
    allreduce_buffers(agg, wgt, group)
    seg, probs = net.finalize(agg, wgt)
    sl = tuple(slicer[1:])
    return seg[sl], probs[(slice(None),) + sl]


def ensemble_mean(softmaxes: Sequence[torch.Tensor]) -> torch.Tensor:
    """a13: np.mean(np.vstack(softmax), 0) over the k checkpoints, on device."""
    acc = torch.zeros_like(softmaxes[0])
    for s in softmaxes:
        acc += s
    return acc / float(len(softmaxes))


def checkpoint_ensemble_refined_label(trainer, checkpoints: Sequence[dict], data, valid_mask=None, voxel_size=(1.0, 1.0, 1.0),
                                      do_mirroring: bool = False):
    """stage-2 label refinement of deepwmh/pipeline/DCNN_multistage.py:317-394 without the NIfTI round trips.
    For each of the k epoch checkpoints: load it (`load_checkpoint_ram`), predict the volume (the reference disables TTA
    here, :334-336), keep the BACKGROUND probability x; accumulate y = 1 - m (1 - x) (`_parallel_softmax_masking`), average
    over k, label = field < 0.5, remove components below 3 mm^3 (`_parallel_ensembling`).
    data: preprocessed [1, X, Y, Z]; valid_mask: [X, Y, Z] array / tensor or None.
    -> (field fp32 [X, Y, Z], label uint8 [X, Y, Z]) as device tensors."""
    net = trainer.network
    k = len(checkpoints)
    if k == 0:
        raise ValueError("checkpoint_ensemble_refined_label: no checkpoints")
    acc = None
    m = None
    if valid_mask is not None:
        m = torch.as_tensor(np.ascontiguousarray(valid_mask) if isinstance(valid_mask, np.ndarray) else valid_mask)
        m = m.to(device=net.device, dtype=torch.float32).contiguous()
    for ck in checkpoints:
        trainer.load_checkpoint_ram(ck, False)
        _, softmax = trainer.predict_preprocessed_data_return_seg_and_softmax(
            data, do_mirroring=do_mirroring, mirror_axes=trainer.data_aug_params["mirror_axes"], use_sliding_window=True,
            step_size=0.5, use_gaussian=True, all_in_gpu=False, mixed_precision=True)
        bg = torch.as_tensor(softmax[0]).to(device=net.device, dtype=torch.float32).contiguous()
        if acc is None:
            acc = torch.zeros_like(bg)
        with torch.cuda.device(net.device):
            net.ensemble_masked_add_(acc, bg, m)
    with torch.cuda.device(net.device):
        label = net.ensemble_refine_(acc, k)
        label = net.remove_3mm_sparks(label, voxel_size)
    return acc, label
