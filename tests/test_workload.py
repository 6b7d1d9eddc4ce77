"""deepwmh_b200/workload.py (product-side synthetic inputs for bench.py) against the oracle's definitions of the same workload."""
import numpy as np
import pytest

import oracle as O
from conftest import small_plans
from deepwmh_b200 import benchmark_plans, workload as W


def test_synthetic_volume_is_the_oracles():
    for shape, seed in (((40, 50, 45), 0), ((33, 20, 27), 5)):
        assert np.array_equal(W.synthetic_flair(shape, seed), O.synthetic_flair(shape, seed))
    v = W.synthetic_flair((40, 50, 45), 1)
    assert v.dtype == np.float32 and v.shape == (1, 40, 50, 45) and (v[0, 0, 0] == 0).all() and v.max() > 150


@pytest.mark.parametrize("plans", [benchmark_plans(), small_plans(), small_plans(patch=(16, 64, 48), pools=((1, 2, 2), (2, 2, 2), (2, 2, 2)),
                                                                              kernels=[[1, 3, 3], [3, 3, 3], [3, 3, 3], [3, 3, 3]])])
def test_flops_and_state_dict_layout_match_the_oracle_network(plans):
    assert W.forward_flops(plans) == O.forward_flops(plans)
    ref = {k: tuple(v.shape) for k, v in O.build_network(plans).state_dict().items()}
    assert dict(W.state_dict_layout(plans)) == ref


def test_benchmark_numbers():
    assert abs(W.forward_flops(benchmark_plans()) / 1e9 - 954.46) < 0.01
    assert sum(int(np.prod(s)) for _, s in W.state_dict_layout(benchmark_plans())) == O.count_parameters(O.build_network(benchmark_plans()))


def test_random_init_statistics_and_loadability():
    plans = small_plans()
    sd = W.random_init_state_dict(plans, 0)
    w = sd["conv_blocks_context.1.blocks.1.conv.weight"]
    fan_in = w.shape[1] * 27
    assert abs(w.std() / np.sqrt(2.0 / (1.0001 * fan_in)) - 1) < 0.05 and abs(w.mean()) < 1e-3
    g = sd["conv_blocks_context.0.blocks.0.instnorm.weight"]
    assert 0.5 <= g.min() and g.max() <= 1.5
    assert not np.array_equal(sd["tu.0.weight"], W.random_init_state_dict(plans, 1)["tu.0.weight"])
    assert np.array_equal(sd["tu.0.weight"], W.random_init_state_dict(plans, 0)["tu.0.weight"])
    import torch
    net = O.build_network(plans)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)      # every key, every shape
