"""Host vs device time of one device-resident assistance round (development aid, not part of the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import dmtcdr_b200
from dmtcdr_b200 import roundloop, engine as E, native

data, dataset, data_split, mats, cfg = bench.build_problem()
R = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, local_epochs=20, device="cuda:0")
R.round0()
for t in range(1, 3):
    R.run_round(t)
R.sync()
for t in range(3, 6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R.run_round(t)
    t1 = time.perf_counter()
    R.sync()
    t2 = time.perf_counter()
    print("round %d: host enqueue %.1f ms, total %.1f ms" % (t, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))
# host pieces
eng = R.eng[0]
t0 = time.perf_counter()
for _ in range(20):
    lay = E.EpochLayout(E.fast_perm_batches(R.n_rows, 500, R.host_gen), eng.d_len, eng.t_len)
print("20 layouts: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter()
flat0 = roundloop.init_flat_params(eng.n_enc, eng.n_dec, 256, 128, "cuda:0", R.gen)
torch.cuda.synchronize()
print("init params: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
# single org alone: 20 epochs
layouts = [E.EpochLayout(E.fast_perm_batches(R.n_rows, 500, R.host_gen), eng.d_len, eng.t_len) for _ in range(20)]
eng.set_round(flat0, R.residual["train"])
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.enqueue_epochs(layouts, list(range(20)), hp=R.hp)
t1 = time.perf_counter()
eng.sync()
t2 = time.perf_counter()
print("one org 20 epochs: host %.1f ms, total %.1f ms (%.1f us/step)" % (1e3 * (t1 - t0), 1e3 * (t2 - t0), 1e6 * (t2 - t0) / 260))
print(eng.h.profile_step(0, 20))
