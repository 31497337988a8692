"""Device-resident assistance rounds at ML1M shape: ms per round and the per-class step profile of one organization.
DMT_DECODER=gather|tc selects the decoder form, DMT_FANOUT=0|1 the step shape (A/B runs).
Usage: python scripts/ab_round.py [rounds] [local_epochs] [world]   (world > 1: rank 0 of an emulated org-sharded run:
only its ceil(18/world) organizations train and predict; no exchange)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop

n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
group = len(sys.argv) > 4 and sys.argv[4] == "group"
data, dataset, data_split, mats, cfg = bench.build_problem()
R = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, local_epochs=epochs, device="cuda:0",
                           rank=0, world=world, group=group)
R.round0()


def one(t):
    R.train_predict(t)
    R.combine()


for t in (1, 2):
    one(t)
R.sync()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(3, 3 + n_rounds):
    one(t)
e1.record()
R.sync()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n_rounds
eng = R.eng[R.my_orgs[0]]
prof = eng.h.profile_step(b=0, reps=20)
print(json.dumps({"decoder": eng.decoder, "fanout": R.fanout, "whole_round": R.whole_round, "group": group, "world": world, "orgs": len(R.my_orgs), "ms_per_round": ms,
                  "step_kernel_ms_org0": {k: round(v, 5) for k, v in prof.items()},
                  "step_sum_us": 1e3 * sum(prof.values())}))
