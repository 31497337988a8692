"""Tensor-core (tcgen05, 3xTF32) kernels against fp64 algebra and against the gather kernels they replace.

The parity mode is passes = 3 (hi/lo TF32 split, three MMAs per k-step): it has to hold the same tolerances as the
fp32 FFMA / SDDMM kernels (loss <= 1e-5 relative, BASELINE.json north_star). passes = 1 is the labelled
reduced-precision mode: the test only pins that it IS a single TF32 pass (error ~1e-3, far above fp32) and not broken.
"""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from oracle import models as om
from golden_io import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import native

    native.load()
    assert native.load().dmt_check_device() == 0
    return native


def cu(x, dtype=None):
    t = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("m,n,k", [(1, 128, 256), (37, 128, 256), (500, 256, 128), (700, 64, 100), (1000, 130, 33)])
def test_dense_layers_tc(nat, act, m, n, k, passes):
    g = torch.Generator().manual_seed(m * n + k)
    X = torch.randn(m, k, generator=g)
    W = torch.randn(n, k, generator=g) / k ** 0.5
    b = torch.randn(n, generator=g) * 0.1
    keep = (torch.rand(m, n, generator=g) < 0.5).to(torch.uint8)
    f = [lambda z: z, torch.tanh][act]
    X64, W64, b64 = X.double(), W.double(), b.double()
    pre = f(X64 @ W64.t() + b64)
    Yr = pre * keep.double() * 2.0
    tol_f, tol_b = (5e-6, 1e-5) if passes == 3 else (5e-3, 5e-3)  # 3xTF32: ~2^-21 per product
    Y, Y_pre = nat.dense_fwd_tc(cu(X), cu(W), cu(b), act, cu(keep), 2.0, passes=passes)
    assert rel_err(Y_pre.cpu(), pre) < tol_f
    assert rel_err(Y.cpu(), Yr) < tol_f
    if passes == 1 and k >= 100 and m >= 37:
        assert rel_err(Y_pre.cpu(), pre) > 2e-5  # really one TF32 pass
    dY = torch.randn(m, n, generator=g)
    dact = [torch.ones_like(pre), 1 - pre ** 2][act]
    dZ = (dY.double() * keep.double() * 2.0 * dact).float()
    dW, db = nat.dense_bwd_w_tc(cu(dZ), cu(X), passes=passes)
    assert rel_err(dW.cpu(), dZ.double().t() @ X64) < tol_b
    assert rel_err(db.cpu(), dZ.double().sum(0)) < 1e-5
    Aprev = torch.tanh(torch.randn(m, k, generator=g))
    dX = nat.dense_bwd_x_tc(cu(dZ), cu(W), cu(Aprev), 1, passes=passes)
    assert rel_err(dX.cpu(), (dZ.double() @ W64) * (1 - Aprev.double() ** 2)) < tol_b


def _rand_csr(g, n_rows, n_cols, density, empty_rows=(), full_rows=()):
    dense = (torch.rand(n_rows, n_cols, generator=g) < density)
    for r in empty_rows:
        dense[r] = False
    for r in full_rows:
        dense[r] = True
    val = torch.randint(1, 6, (n_rows, n_cols), generator=g).float() * dense
    m = csr_matrix(val.numpy())
    m.sort_indices()
    return m


@pytest.mark.parametrize("loss", ["explicit", "implicit"])
@pytest.mark.parametrize("n_rows,n_dec,H,density", [(70, 150, 256, 0.3), (300, 700, 256, 0.15), (129, 129, 128, 0.5),
                                                    (500, 3706, 256, 0.04)])
def test_decoder_tc_vs_algebra_and_gather(nat, loss, n_rows, n_dec, H, density):
    """dmt_ae_decoder_tc == dense fp64 algebra of src/models/ae.py:135-156 (+ autograd) == the SDDMM kernel."""
    g = torch.Generator().manual_seed(n_rows + n_dec)
    T = _rand_csr(g, n_rows, n_dec, density, empty_rows=(4, 9), full_rows=(11,))
    if loss == "implicit":
        T.data = (T.data >= 3.5).astype(np.float32)
    A3 = torch.tanh(torch.randn(n_rows, H, generator=g))
    W4 = torch.randn(n_dec, H, generator=g) * 0.1
    b4 = torch.randn(n_dec, generator=g) * 0.1
    rows = torch.arange(n_rows, dtype=torch.int32)
    rt = torch.from_numpy(np.repeat(np.arange(n_rows), np.diff(T.indptr))).long()
    ct = torch.from_numpy(T.indices).long()
    y = torch.from_numpy(T.data).double()
    A3r, W4r, b4r = (A3.double().requires_grad_(True), W4.double().requires_grad_(True),
                     b4.double().requires_grad_(True))
    pred_ref = (A3r[rt] * W4r[ct]).sum(-1) + b4r[ct]
    loss_ref = om.loss_fn(pred_ref, y, loss)
    gout_ref, dA3_ref, dW4_ref, db4_ref = torch.autograd.grad(loss_ref, [pred_ref, A3r, W4r, b4r])
    args = (cu(rows), cu(T.indptr, torch.int32), cu(T.indices, torch.int32), cu(T.data), cu(A3), cu(W4), cu(b4),
            nat.LOSS_KIND[loss], T.nnz, True)
    pred, gout, dz3, dW4, db4, loss_rows, n_t = nat.ae_decoder_tc(*args)
    assert int(n_t) == T.nnz
    assert rel_err(pred.cpu(), pred_ref.detach()) < 3e-6
    assert rel_err(gout.cpu(), gout_ref) < 1e-5
    assert abs(float(loss_rows.double().sum()) / T.nnz - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert float(loss_rows[1:].abs().sum()) == 0.0
    assert rel_err(dz3.cpu(), dA3_ref * (1 - A3.double() ** 2)) < 1e-5
    assert rel_err(dW4.cpu(), dW4_ref) < 1e-5
    assert rel_err(db4.cpu(), db4_ref) < 1e-5
    # the gather kernel it replaces
    pred_g, gout_g, dz3_g, loss_rows_g, _ = nat.ae_decoder_fwd(*args)
    assert rel_err(pred.cpu(), pred_g.cpu()) < 3e-6
    assert rel_err(dz3.cpu(), dz3_g.cpu()) < 1e-5
    # run-to-run determinism (no atomics, fixed summation orders)
    again = nat.ae_decoder_tc(*args)
    for a, b in zip((pred, gout, dz3, dW4, db4, loss_rows), again):
        assert torch.equal(a, b)
    # eval mode: predictions only
    pred_e, *_ = nat.ae_decoder_tc(cu(rows), cu(T.indptr, torch.int32), cu(T.indices, torch.int32), None, cu(A3),
                                   cu(W4), cu(b4), nat.LOSS_KIND[loss], T.nnz, False)
    assert torch.equal(pred_e, pred)


def test_decoder_tc_row_subset_and_single_pass(nat):
    """A batch that is a subset of the CSR's rows in arbitrary order (the engine's case), and the 1-pass mode."""
    g = torch.Generator().manual_seed(5)
    n_all, n_dec, H = 400, 300, 256
    T = _rand_csr(g, n_all, n_dec, 0.2, empty_rows=(0, 17))
    rows = torch.randperm(n_all, generator=g)[:200].to(torch.int32)
    A3 = torch.tanh(torch.randn(len(rows), H, generator=g))
    W4 = torch.randn(n_dec, H, generator=g) * 0.1
    b4 = torch.randn(n_dec, generator=g) * 0.1
    args = (cu(rows), cu(T.indptr, torch.int32), cu(T.indices, torch.int32), cu(T.data), cu(A3), cu(W4), cu(b4), 0,
            T.nnz, True)
    pred, gout, dz3, dW4, db4, loss_rows, n_t = nat.ae_decoder_tc(*args)
    pred_g, gout_g, dz3_g, loss_rows_g, n_t_g = nat.ae_decoder_fwd(*args)
    assert int(n_t) == int(n_t_g)
    assert rel_err(pred.cpu(), pred_g.cpu()) < 3e-6
    assert rel_err(gout.cpu(), gout_g.cpu()) < 1e-5
    assert rel_err(dz3.cpu(), dz3_g.cpu()) < 1e-5
    assert abs(float(loss_rows.sum()) - float(loss_rows_g.sum())) <= 1e-5 * abs(float(loss_rows_g.sum()))
    # dW4 / db4 from the gathered g (fp64)
    Gd = torch.zeros(len(rows), n_dec, dtype=torch.float64)
    ip = T.indptr
    for j, u in enumerate(rows.tolist()):
        sl = slice(ip[u], ip[u + 1])
        Gd[j, T.indices[sl]] = gout_g.cpu().double()[sl]
    assert rel_err(dW4.cpu(), Gd.t() @ A3.double()) < 1e-5
    assert rel_err(db4.cpu(), Gd.sum(0)) < 1e-5
    p1, g1, dz1, dW1, *_ = nat.ae_decoder_tc(*args, passes=1)
    e = rel_err(p1.cpu(), pred_g.cpu())
    assert 2e-5 < e < 5e-3
