"""cProfile of the drop-in e2e rounds (host-side overhead hunt): python scripts/profile_e2e.py [rounds]"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import runner, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
data = synth.make_rating_data("ML1M", seed=0)
marks = []
def on_round(t):
    torch.cuda.synchronize()
    marks.append(time.perf_counter())
pr = cProfile.Profile()
use_prof = os.environ.get("E2E_PROFILE", "1") == "1"
def run():
    return runner.run_assist_experiment(data, bench.CONTROL, seed=0, local_epochs=20, rounds=n, rng="device", on_round=on_round)
if use_prof:
    pr.enable()
res = run()
pr.disable()
print("round ms:", [round(1e3 * (b - a), 1) for a, b in zip(marks[:-1], marks[1:])])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
