"""Do kernels of two independent contexts overlap on the GPU?  Runs forward_patches on two contexts from two
host threads / streams and compares with running them back to back."""
import os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O
import deepwmh_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
plans = deepwmh_b200.benchmark_plans()
net = O.build_benchmark_network(0, plans)
trs = []
for i in range(2):
    tr = deepwmh_b200.nnUNetTrainerV2(plans, device=0, max_batch=n, lanes=1)
    tr.load_checkpoint_ram({"state_dict": net.state_dict()}, False)
    trs.append(tr)
x = torch.randn(n, 1, 128, 128, 128, generator=torch.Generator().manual_seed(0)).cuda()
streams = [torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)]

def run(i, reps):
    with torch.cuda.stream(streams[i]):
        for _ in range(reps):
            trs[i].network.forward_patches(x)

for i in range(2):
    run(i, 1)
torch.cuda.synchronize()
t0 = time.time(); run(0, 3); run(1, 3); torch.cuda.synchronize(); t_seq = time.time() - t0
t0 = time.time()
th = [threading.Thread(target=run, args=(i, 3)) for i in range(2)]
[t.start() for t in th]; [t.join() for t in th]
torch.cuda.synchronize(); t_par = time.time() - t0
print("n=%d: sequential %.1f ms, two streams %.1f ms (ratio %.2f)" % (n, t_seq * 1e3, t_par * 1e3, t_seq / t_par))
