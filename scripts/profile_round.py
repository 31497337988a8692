"""Two short device-resident assistance rounds (2 local epochs) for ncu launch lists / captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
group = len(sys.argv) > 2 and sys.argv[2] == "group"
data, dataset, data_split, mats, cfg = bench.build_problem()
R = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, local_epochs=epochs, device="cuda:0",
                           group=group)
R.round0()
for t in (1, 2):
    R.run_round(t)
R.sync()
print("ok")
