"""Scaled-shape run of the config-5 family (BASELINE.json configs[4]: 1M users x 500K items, 1B ratings, 64
organizations over 8 GPUs), reduced so that one GPU call of a few minutes covers it:

    python scripts/scaled_round.py [M N NNZ K WORLD EPOCHS]      defaults: 400000 200000 100000000 64 8 1

builds M x N synthetic ratings (Zipf item popularity), splits the items into K random organizations
(`random-K`, src/data.py:231-235) and runs assistance rounds for RANK 0 of a WORLD-rank org-sharded job on this GPU:
its K / WORLD organizations train (EPOCHS local epochs, batches of 500 users, per-epoch plans) and predict, then the
combine runs over all K rows of the prediction matrix exactly as on every rank of the real job (the other ranks' rows
stay zero: same bytes, same kernels; only the NCCL all-gather itself is not exercised here - bench.py --gpus N does
that at ML1M shape). Prints one JSON object: round time, rating-visits/s of this rank, the per-kernel-class step times
of one organization with their algorithmic bytes against the measured HBM peak, and the round's aggregate rate.
Launched under torchrun (one process per GPU) it runs the REAL org-sharded job instead: WORLD = the launcher's world
size, every rank trains its share, the prediction rows are exchanged with the in-place NCCL all-gather and the round
time is the max over ranks (rank 0 prints). Not part of the bench contract (bench.py is the ML1M headline)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import dmtcdr_b200  # noqa: F401
from dmtcdr_b200 import dist as D
from dmtcdr_b200 import engine as E
from dmtcdr_b200 import roundloop, synth

a = [int(x) for x in sys.argv[1:]]
M, N, NNZ, K, WORLD, EPOCHS = (a + [400_000, 200_000, 100_000_000, 64, 8, 1][len(a):])[:6]
BS = 500
sys.stdout.flush()
_saved_stdout = os.dup(1)
os.dup2(2, 1)  # libraries (NCCL's banner) print to the C-level stdout: keep it for the one JSON object
RANK, REAL_WORLD, LOCAL = D.init_from_env()
REAL = REAL_WORLD > 1
if REAL:
    WORLD = REAL_WORLD
torch.cuda.set_device(LOCAL)
DEV = "cuda:{}".format(LOCAL)
t0 = time.time()
data = synth.make_scaled_data_device(M, N, NNZ, 8, seed=0, device=DEV)
gen_s = time.time() - t0
torch.manual_seed(0)
chunks = list(torch.randperm(N).split(N // K))
split = [c.numpy() for c in chunks[:K - 1]] + [torch.cat(chunks[K - 1:]).numpy()]
mats = {"train": (data.train, data.train), "test": (data.train, data.test)}
t0 = time.time()
R = roundloop.AssistRounds(mats, split, "explicit", BS, local_epochs=EPOCHS, device=DEV, rank=RANK if REAL else 0,
                           world=WORLD, whole_round=False)
R.round0()
setup_s = time.time() - t0
exchange = (lambda O: D.exchange_outputs(R.state.O_full, R.chunk, RANK, WORLD)) if REAL else None
R.run_round(1, exchange)
R.sync()
ts = []
for t in (2, 3):
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    R.run_round(t, exchange)
    e1.record()
    R.sync()
    ts.append(D.max_over_ranks(e0.elapsed_time(e1) * 1e-3, DEV))
sec = min(ts)
org = R.my_orgs[0]
eng = R.eng[org]
prof = eng.h.profile_step(b=0, reps=10)
hbm, src = bench.measured_peaks()
n_params = eng.h.n_params
n_tr, n_te = data.train.nnz, data.test.nnz
nb = -(-M // BS)
t_batch, d_batch = n_tr / nb, float(eng.d_len.sum()) / nb
by_class = bench.step_algorithmic_bytes(t_batch, d_batch, min(N, int(t_batch)), min(eng.n_enc, int(d_batch)), BS, n_params)
classes = {k: {"ms": v, "algorithmic_bytes": by_class[k], "GBps": by_class[k] / (v * 1e-3) / 1e9,
               "frac_of_hbm_peak": by_class[k] / (v * 1e-3) / 1e9 / hbm} for k, v in prof.items() if k in by_class}
n_local = len(R.my_orgs)
visits_rank = n_local * (EPOCHS * n_tr + n_tr + n_te)
agg = E.bytes_per_round(n_local, n_tr, n_te, M, [len(split[k]) for k in R.my_orgs], N, EPOCHS, BS)
agg += (n_tr + n_te) * (4 * K + 24) - (n_tr + n_te) * (4 * n_local + 24)  # the combine reads all K rows
if REAL:
    D.barrier()
    import torch.distributed as tdist
    tdist.destroy_process_group()
    if RANK != 0:
        sys.exit(0)
out = {"mode": "real ranks (NCCL all-gather, max over ranks)" if REAL else "rank 0 of an emulated job",
       "shape": {"users": M, "items": N, "train": int(n_tr), "test": int(n_te)}, "organizations": K, "world": WORLD,
       "organizations_on_this_rank": n_local, "local_epochs": EPOCHS, "batch_rows": BS,
       "host_seconds": {"generate": gen_s, "setup_and_round0": setup_s},
       "round_ms": 1e3 * sec, "rating_visits_per_s_this_rank": visits_rank / sec,
       "rating_visits_per_s_job" if REAL else "rating_visits_per_s_job_if_all_ranks_match":
           K * (EPOCHS * n_tr + n_tr + n_te) / sec,
       "n_params_per_org": int(n_params), "step_classes": classes, "step_sum_us": 1e3 * sum(prof.values()),
       "round_aggregate": {"algorithmic_bytes": agg, "GBps": agg / sec / 1e9, "frac_of_hbm_peak": agg / sec / 1e9 / hbm},
       "hbm_peak_GBps": hbm, "peak_source": src,
       "device_memory_used_GB": (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9}
sys.stdout.flush()
os.dup2(_saved_stdout, 1)
print(json.dumps(out), flush=True)
