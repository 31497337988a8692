import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop
data, dataset, data_split, mats, cfg = bench.build_problem()
for world in (18, 9, 4, 2, 1):
    R = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, local_epochs=20, rank=0, world=world, device="cuda:0")
    R.round0()
    R.run_round(1); R.sync()
    ts = []
    for t in range(2, 5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        R.run_round(t); R.sync()
        ts.append(time.perf_counter() - t0)
    n = len(R.my_orgs)
    print("orgs %2d: %.1f ms/round  (%.1f ms per org)" % (n, 1e3 * min(ts), 1e3 * min(ts) / n), flush=True)
    R.close(); del R
    torch.cuda.empty_cache()
