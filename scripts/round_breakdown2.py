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
st = R.state
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0) + time.perf_counter() - t0
torch.cuda.synchronize()
a0 = torch.cuda.memory_stats()["num_device_alloc"]
tt = time.perf_counter()
t0 = time.perf_counter()
for k in R.splits:
    st.residual(R.F[k], k, R.clamp, out=R.residual[k])
tick("residual", t0)
for org in R.my_orgs:
    eng = R.eng[org]
    t0 = time.perf_counter(); flat0 = roundloop.init_flat_params(eng.n_enc, eng.n_dec, 256, 128, "cuda:0", R.gen); tick("init", t0)
    t0 = time.perf_counter(); eng.set_round(flat0, R.residual["train"]); tick("set_round", t0)
    t0 = time.perf_counter(); layouts = [E.EpochLayout(E.fast_perm_batches(R.n_rows, 500, R.host_gen), eng.d_len, eng.t_len) for _ in range(20)]; tick("layouts", t0)
    t0 = time.perf_counter(); lb = torch.zeros(sum(len(l.active) for l in layouts), device="cuda:0"); tick("zeros", t0)
    t0 = time.perf_counter(); eng.enqueue_epochs(layouts, list(range(20)), hp=R.hp, loss_out=lb); tick("enqueue", t0)
for org in R.my_orgs:
    eng = R.eng[org]
    t0 = time.perf_counter()
    eng.predict(R.org_data[org], st.y["train"], st.o_row("train", org))
    eng.predict(R.org_test_data[org], st.y["test"], st.o_row("test", org))
    tick("predict", t0)
t0 = time.perf_counter()
for org in R.my_orgs:
    R.eng[org].h.signal_current()
tick("signal", t0)
t0 = time.perf_counter(); F_next, fitted = st.update(R.F, 0.1); tick("update", t0)
t0 = time.perf_counter(); torch.cuda.synchronize(); tick("final_sync", t0)
print("total %.1f ms" % (1e3 * (time.perf_counter() - tt)), {k: round(1e3 * v, 2) for k, v in T.items()})
print("device allocs during round:", torch.cuda.memory_stats()["num_device_alloc"] - a0)
