"""One ML1M-shaped batch (500 rows x 3706 items, ~69K targets) through the decoder's two forms via the stateless C-ABI
calls: CUDA-event time per call, and a short loop for ncu (`-k regex:dec_|ae_decoder|segment`)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.sparse import csr_matrix
import dmtcdr_b200
from dmtcdr_b200 import native as nat

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
n_rows, n_dec, H = 500, 3706, 256
pop = 1.0 / np.arange(1, n_dec + 1) ** 0.8
pop = rng.permutation(pop / pop.sum())
lens = np.clip(rng.lognormal(4.5, 0.9, n_rows).astype(int), 18, 2000)
rows_i, cols_i = [], []
for r in range(n_rows):
    c = rng.choice(n_dec, size=lens[r], replace=False, p=pop)
    rows_i.append(np.full(len(c), r)); cols_i.append(c)
rows_i, cols_i = np.concatenate(rows_i), np.concatenate(cols_i)
T = csr_matrix((rng.normal(size=len(rows_i)).astype(np.float32), (rows_i, cols_i)), shape=(n_rows, n_dec))
T.sort_indices()
dev = "cuda"
cu = lambda x, dt=None: torch.as_tensor(np.asarray(x)).to(dt if dt else torch.as_tensor(np.asarray(x)).dtype).to(dev).contiguous()
g = torch.Generator().manual_seed(0)
A3 = torch.tanh(torch.randn(n_rows, H, generator=g)).to(dev)
W4 = (torch.randn(n_dec, H, generator=g) * 0.05).to(dev)
b4 = torch.zeros(n_dec, device=dev)
rows = torch.arange(n_rows, dtype=torch.int32, device=dev)
args = (rows, cu(T.indptr, torch.int32), cu(T.indices, torch.int32), cu(T.data), A3, W4, b4, 0, T.nnz, True)

def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# note: the python wrappers allocate outputs per call; the allocator caches, so this adds host time only
t_tc = timed(lambda: nat.ae_decoder_tc(*args, passes=passes))
t_g = timed(lambda: nat.ae_decoder_fwd(*args))
t_e = timed(lambda: nat.ae_decoder_tc(*args[:3], None, *args[4:9], False, passes=passes))
flops = 2.0 * n_rows * n_dec * H
print(json.dumps({"nnz": int(T.nnz), "passes": passes, "tc_train_ms(D1+D2+finish+D3+tab)": t_tc,
                  "gather_fwd_dz3_ms(no dW4)": t_g, "tc_eval_fwd_ms(D1)": t_e,
                  "D1_algorithmic_TFLOPs": flops / (t_e * 1e-3) / 1e12}))
