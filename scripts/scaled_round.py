"""Scaled-shape sanity/roofline run (config-5 family, reduced to fit a short GPU call):
   100 000 users x 50 000 items, ~20 M ratings, 8 random organizations, batch 500 rows, 1 local epoch per round.
   Prints one JSON object with the round time, the per-kernel-class step times of organization 0 and the
   HBM fractions of the streaming kernels. Not part of the bench contract (bench.py is the ML1M headline)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop, synth, native

M, N, NNZ, K = 100_000, 50_000, 20_000_000, 8
t0 = time.time()
data = synth.make_scaled_data(M, N, NNZ, 8, seed=0)
gen_s = time.time() - t0
torch.manual_seed(0)
chunks = list(torch.randperm(N).split(N // K))
split = [c.numpy() for c in chunks[:K - 1]] + [torch.cat(chunks[K - 1:]).numpy()]
mats = {"train": (data.train, data.train), "test": (data.train, data.test)}
R = roundloop.AssistRounds(mats, split, "explicit", 500, local_epochs=1, device="cuda:0")
R.round0()
R.run_round(1); R.sync()
ts = []
for t in (2, 3):
    torch.cuda.synchronize(); a = time.perf_counter()
    R.run_round(t); R.sync()
    ts.append(time.perf_counter() - a)
eng = R.eng[0]
prof = eng.h.profile_step(b=0, reps=10)
hbm, src = bench.measured_peaks()
n_params = eng.h.n_params
t_batch = float(np.mean([eng.t_len[r:r + 500].sum() for r in range(0, 5000, 500)]))
visits = K * (1 * data.train.nnz + data.train.nnz + data.test.nnz)
out = {"shape": [M, N, int(data.train.nnz), int(data.test.nnz)], "orgs": K, "gen_seconds": gen_s,
       "round_ms": 1e3 * min(ts), "rating_visits_per_s": visits / min(ts), "step_kernel_ms": prof,
       "n_params_per_org": int(n_params),
       "adam": {"bytes": n_params * 32, "GBps": n_params * 32 / (prof["clip_adam"] * 1e-3) / 1e9,
                "frac_of_hbm_peak": n_params * 32 / (prof["clip_adam"] * 1e-3) / 1e9 / hbm},
       "decoder": {"targets_per_batch": t_batch, "bytes": t_batch * (4 * 256 + 20),
                   "GBps": t_batch * (4 * 256 + 20) / (prof["decoder_loss_dz3"] * 1e-3) / 1e9},
       "grad_norm": {"GBps": n_params * 4 / (prof["grad_norm"] * 1e-3) / 1e9}, "hbm_peak": hbm}
print(json.dumps(out))
