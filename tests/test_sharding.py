"""CPU: the host logic of the multi-GPU modes, world_size 2 over gloo.
The per-rank tile accumulation is played by the oracle here (no GPU); what is under test is the
partitioning, the one collective (reduce to rank 0 / all-reduce of agg, weights computed locally; and round 1's
two-buffer all-reduce) and that the result is independent of the shard count."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from conftest import small_plans
from deepwmh_b200.parallel import ensemble_mean, reduce_aggregation, shard_cohort, shard_tiles


def test_shard_cohort_partitions():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            parts = [shard_cohort(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_shard_tiles_contiguous_cover():
    for n in (1, 12, 196, 197):
        for w in (1, 2, 4, 8):
            prev = 0
            for r in range(w):
                b, e = shard_tiles(n, r, w)
                assert b == prev and e >= b
                prev = e
            assert prev == n


def test_ensemble_mean():
    a, b = torch.rand(2, 3, 4, 5), torch.rand(2, 3, 4, 5)
    assert torch.allclose(ensemble_mean([a, b]), torch.from_numpy(np.mean(np.stack([a.numpy(), b.numpy()]), 0)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _tile_range_buffers(net, x, patch, b, e):
    """Oracle-side stand-in for dwmh_predict_3d(tile_begin=b, tile_end=e)."""
    steps = O.compute_steps_for_sliding_window(patch, x.shape[1:], 0.5)
    g = O.get_gaussian(patch)
    agg = np.zeros((2,) + x.shape[1:], np.float32); wgt = np.zeros(x.shape[1:], np.float32)
    lin = 0
    for lx in steps[0]:
        for ly in steps[1]:
            for lz in steps[2]:
                if b <= lin < e:
                    sl = (slice(lx, lx + patch[0]), slice(ly, ly + patch[1]), slice(lz, lz + patch[2]))
                    t = torch.from_numpy(np.ascontiguousarray(x[(None, slice(None)) + sl]))
                    agg[(slice(None),) + sl] += O.mirror_and_predict(net, t, (0, 1, 2), False, torch.from_numpy(g))[0].numpy()
                    wgt[sl] += g
                lin += 1
    return agg, wgt, lin


def _full_weight_map(shape, patch):
    """Host stand-in for dwmh_weight_map: the importance map of every tile, added in tile order."""
    steps = O.compute_steps_for_sliding_window(patch, shape, 0.5)
    g = O.get_gaussian(patch)
    wgt = np.zeros(shape, np.float32)
    for lx in steps[0]:
        for ly in steps[1]:
            for lz in steps[2]:
                wgt[lx:lx + patch[0], ly:ly + patch[1], lz:lz + patch[2]] += g
    return wgt


def _worker(rank, world, port, out, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    plans = small_plans(patch=(16, 16, 16), pools=((2, 2, 2),) * 2)
    net = O.build_benchmark_network(0, plans)
    x = np.random.default_rng(0).normal(size=(1, 24, 28, 20)).astype(np.float32)
    steps = O.compute_steps_for_sliding_window((16,) * 3, x.shape[1:], 0.5)
    n_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    b, e = shard_tiles(n_tiles, rank, world)
    agg, wgt, _ = _tile_range_buffers(net, x, (16, 16, 16), b, e)
    agg_t, wgt_t = torch.from_numpy(agg), torch.from_numpy(wgt)
    complete = reduce_aggregation(agg_t, wgt_t if mode == "allreduce2" else None, mode)
    assert complete == (rank == 0 or mode != "reduce")
    if mode != "allreduce2":
        wgt_t = torch.from_numpy(_full_weight_map(x.shape[1:], (16, 16, 16)))      # computed locally, never reduced
    # cohort: every rank lists its subjects; gathered on host
    mine = shard_cohort(5, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        np.savez(out, agg=agg_t.numpy(), wgt=wgt_t.numpy(), cohort=np.array(sorted(sum(gathered, []))))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["reduce", "allreduce", "allreduce2"])
def test_tile_sharded_collective_world2(tmp_path, mode):
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, _free_port(), out, mode), nprocs=2, join=True)
    r = np.load(out)
    plans = small_plans(patch=(16, 16, 16), pools=((2, 2, 2),) * 2)
    net = O.build_benchmark_network(0, plans)
    x = np.random.default_rng(0).normal(size=(1, 24, 28, 20)).astype(np.float32)
    agg, nb = O.predict_3D_tiled(net, x, 0.5, False, (0, 1, 2), (16, 16, 16), True, return_buffers=True)
    assert np.allclose(r["agg"], agg, rtol=1e-5, atol=1e-7)      # fp32 summation order differs across shard counts
    if mode == "allreduce2":
        assert np.allclose(r["wgt"], nb[0], rtol=1e-6)
    else:
        assert np.array_equal(r["wgt"], nb[0])                   # tile-ordered local sum: bit-identical to the single run
    assert list(r["cohort"]) == [0, 1, 2, 3, 4]


def test_reduce_mode_is_validated():
    with pytest.raises(ValueError):
        reduce_aggregation(torch.zeros(2, 2, 2, 2), None, "ring")
