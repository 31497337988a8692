"""Device-resident organization engine (dmt_org_*) against the oracle's Organization.train / predict.

Same parameters, same batches, same dropout keep-masks -> same per-batch losses (<= 1e-5 relative), same
parameters after training (Adam-amplified rounding, <= 5e-4 of max |w|) and same predictions (<= 1e-4).
"""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from oracle import replay, train
from golden_io import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import native

    native.load()
    return native


def cu(x, dtype=None):
    t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def rand_csr(rng, n_rows, n_cols, density, empty_rows=(), values="rating"):
    mask = rng.random((n_rows, n_cols)) < density
    for r in empty_rows:
        mask[r] = False
    if values == "rating":
        val = rng.integers(1, 6, size=(n_rows, n_cols)).astype(np.float32)
    else:
        val = rng.normal(size=(n_rows, n_cols)).astype(np.float32)
        val[val == 0] = 0.5
    m = csr_matrix(val * mask)
    m.sort_indices()
    return m


def flat_params(p):
    """Engine layout: W1t b1 W2 b2 W3 b3 W4 b4 (include/dmt_b200.h, dmt_org_num_params)."""
    return torch.cat([p["encoder_linear.weight"].t().contiguous().view(-1), p["encoder_linear.bias"],
                      p["encoder.blocks.0.weight"].view(-1), p["encoder.blocks.0.bias"],
                      p["decoder.blocks.0.weight"].view(-1), p["decoder.blocks.0.bias"],
                      p["decoder_linear.weight"].view(-1), p["decoder_linear.bias"]])


def unflat_params(flat, n_enc, n_dec, H1=256, H2=128):
    out = {}
    o = 0

    def take(n, shape):
        nonlocal o
        t = flat[o:o + n].view(shape)
        o += n
        return t

    out["encoder_linear.weight"] = take(n_enc * H1, (n_enc, H1)).t().contiguous()
    out["encoder_linear.bias"] = take(H1, (H1,))
    out["encoder.blocks.0.weight"] = take(H2 * H1, (H2, H1))
    out["encoder.blocks.0.bias"] = take(H2, (H2,))
    out["decoder.blocks.0.weight"] = take(H1 * H2, (H1, H2))
    out["decoder.blocks.0.bias"] = take(H1, (H1,))
    out["decoder_linear.weight"] = take(n_dec * H1, (n_dec, H1))
    out["decoder_linear.bias"] = take(n_dec, (n_dec,))
    return out


@pytest.mark.parametrize("decoder", ["gather", "tc"])
@pytest.mark.parametrize("seed,bs,n_rows,n_dec", [(0, 50, 130, 100), (1, 64, 130, 100), (2, 150, 400, 300)])
def test_train_epochs_and_predict(nat, seed, bs, n_rows, n_dec, decoder):
    rng = np.random.default_rng(seed)
    n_enc = 30
    # rows 7,8 have targets but no data; row 9 has nothing; rows 100..129 (a whole batch for bs=50 after sorting the
    # batch list below) have no data -> that batch must be skipped without an optimizer step
    D = rand_csr(rng, n_rows, n_enc, 0.15, empty_rows=(7, 8, 9))
    T = rand_csr(rng, n_rows, n_dec, 0.25, empty_rows=(9,), values="normal")
    torch.manual_seed(seed)
    p0 = replay.init_ae_params(n_enc, n_dec)
    n_epochs = 3
    # batches: fixed permutations; one batch made only of no-data rows
    nodata = [7, 8]
    epoch_batches = []
    for e in range(n_epochs):
        perm = [r for r in rng.permutation(n_rows) if r not in nodata]
        batches = [np.array(perm[s:s + bs]) for s in range(0, len(perm), bs)]
        batches.insert(1, np.array(nodata))
        epoch_batches.append(batches)
    masks = []
    keep_epochs = []
    for batches in epoch_batches:
        ke = []
        for rows in batches:
            rows_eff = [r for r in sorted(rows) if D.indptr[r + 1] > D.indptr[r] or T.indptr[r + 1] > T.indptr[r]]
            k = torch.from_numpy((rng.random((len(rows_eff), 128)) < 0.5).astype(np.float32))
            ke.append(k)
            if any(D.indptr[r + 1] > D.indptr[r] for r in rows):
                masks.append(k)
        keep_epochs.append(ke)
    ref_p, ref_losses = train.train_org_ae(p0, D, T, "user", "explicit", epoch_batches, masks)
    # ---- engine
    d_csr = (cu(D.indptr, torch.int32), cu(D.indices, torch.int32), cu(D.data))
    t_csr = (cu(T.indptr, torch.int32), cu(T.indices, torch.int32))
    org = nat.Org(n_rows, n_enc, n_dec, 256, 128, d_csr, t_csr, bs, 0)
    org.set_decoder_mode(decoder)  # "tc": the decoder's last layer on tcgen05 (3xTF32), same tolerances
    org.wait_current()
    org.set_params(flat_params(p0).cuda())
    tval = cu(T.data)
    org.set_target(tval)
    got_losses = []
    for e, batches in enumerate(epoch_batches):
        rows_l, off = [], [0]
        for rows in batches:
            rows_eff = [r for r in sorted(rows) if D.indptr[r + 1] > D.indptr[r] or T.indptr[r + 1] > T.indptr[r]]
            rows_l += rows_eff
            off.append(len(rows_l))
        rows_a = np.array(rows_l, np.int64)
        n_t = int((T.indptr[rows_a + 1] - T.indptr[rows_a]).sum())
        n_d = int((D.indptr[rows_a + 1] - D.indptr[rows_a]).sum())
        keep = torch.cat(keep_epochs[e]).to(torch.uint8).cuda()
        loss = torch.zeros(len(batches), device="cuda")
        org.wait_current()
        org.train_epoch(cu(rows_a, torch.int32), cu(off, torch.int32), n_t, n_d, keep=keep, epoch_loss=loss)
        org.sync()
        active = [any(D.indptr[r + 1] > D.indptr[r] for r in rows) for rows in batches]
        got_losses += [float(l) for l, a in zip(loss.cpu(), active) if a]
    assert len(got_losses) == len(ref_losses)
    assert rel_err(got_losses, ref_losses) < 1e-5
    flat = org.get_params()
    org.sync()
    got_p = unflat_params(flat.cpu(), n_enc, n_dec)
    # Adam's m/sqrt(v) amplifies rounding of near-zero gradients; 3xTF32 products round ~4x coarser than FFMA fp32
    tol_p = 5e-4 if decoder == "gather" else 1e-3
    for k, v in ref_p.items():
        assert rel_err(got_p[k], v) < tol_p, k
    # ---- predict (all rows at once) vs the oracle's batched eval forward
    T2 = rand_csr(rng, n_rows, n_dec, 0.1, values="normal")
    ref_pred = train.predict_org_ae(ref_p, D, T2, "user", "explicit", bs)
    out = torch.zeros(T2.nnz, device="cuda")
    t2 = (cu(T2.indptr, torch.int32), cu(T2.indices, torch.int32))
    org.set_params(flat_params(ref_p).cuda())
    org.predict(d_csr, t2, n_rows, out)
    org.sync()
    assert rel_err(out.cpu(), ref_pred) < 2e-5
    org.close()


def test_device_dropout_is_a_fair_coin(nat):
    """Without explicit keep-masks the engine draws its own: training must still run and reduce the loss."""
    rng = np.random.default_rng(3)
    n_rows, n_enc, n_dec, bs = 200, 40, 120, 100
    D = rand_csr(rng, n_rows, n_enc, 0.2)
    T = rand_csr(rng, n_rows, n_dec, 0.3, values="normal")
    torch.manual_seed(0)
    p0 = replay.init_ae_params(n_enc, n_dec)
    d_csr = (cu(D.indptr, torch.int32), cu(D.indices, torch.int32), cu(D.data))
    t_csr = (cu(T.indptr, torch.int32), cu(T.indices, torch.int32))
    org = nat.Org(n_rows, n_enc, n_dec, 256, 128, d_csr, t_csr, bs, 0)
    org.wait_current()
    org.set_params(flat_params(p0).cuda())
    tval = cu(T.data)
    org.set_target(tval)
    rows = cu(np.arange(n_rows), torch.int32)
    off = cu([0, 100, 200], torch.int32)
    first = last = None
    for e in range(30):
        loss = torch.zeros(2, device="cuda")
        org.train_epoch(rows, off, T.nnz, D.nnz, seed=123 + e, epoch_loss=loss)
        org.sync()
        if e == 0:
            first = float(loss.mean())
        last = float(loss.mean())
    assert np.isfinite(last) and last < first
    org.close()
