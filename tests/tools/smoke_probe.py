"""GPU: candidate configurations for __graft_entry__.smoke() -- small enough for the CPU oracle, large enough that the three
gate numbers are stable.  Prints the parity report of each.  Test infrastructure (imports oracle/)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402
import deepwmh_b200  # noqa: E402

for patch, pools, shape, seed in [((32, 32, 32), 3, (40, 50, 45), 0), ((64, 64, 64), 4, (96, 112, 96), 0), ((64, 64, 64), 4, (96, 112, 96), 1),
                                   ((64, 64, 64), 4, (80, 100, 90), 0), ((48, 48, 48), 3, (80, 96, 72), 0)]:
    plans = deepwmh_b200.benchmark_plans(patch_size=patch, num_pool=pools)
    net = O.build_benchmark_network(0, plans)
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    raw = O.synthetic_flair(shape, seed=seed)
    data = raw.copy()
    data[0] = O.zscore_nnunet(raw[0], np.where(raw[0] != 0, 0, -1), True)
    t0 = time.time()
    seg_ref, p_ref = O.OracleTrainer(plans, net).predict_preprocessed_data_return_seg_and_softmax(data)
    t_cpu = time.time() - t0
    seg, p = tr.predict_raw_volume_host(np.ascontiguousarray(raw[0]))
    print(patch, pools, shape, seed, "cpu oracle %.1f s" % t_cpu, O.parity_report(seg_ref, p_ref, seg, p), flush=True)
    tr.network.close()
