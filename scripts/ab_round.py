"""Device-resident assistance rounds at ML1M shape: ms per round and the per-class step profile of one organization.
DMT_DECODER=gather|tc selects the decoder form (A/B runs). Usage: python scripts/ab_round.py [rounds] [local_epochs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop

n_rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
data, dataset, data_split, mats, cfg = bench.build_problem()
R = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, local_epochs=epochs, device="cuda:0")
R.round0()
for t in (1, 2):
    R.run_round(t)
R.sync()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(3, 3 + n_rounds):
    R.run_round(t)
e1.record()
R.sync()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n_rounds
eng = R.eng[R.my_orgs[0]]
prof = eng.h.profile_step(b=0, reps=20)
print(json.dumps({"decoder": eng.decoder, "ms_per_round": ms, "visits_per_s": R.rating_visits_per_round() / (ms / 1e3),
                  "step_kernel_ms_org0": prof, "step_sum_us": 1e3 * sum(prof.values())}))
