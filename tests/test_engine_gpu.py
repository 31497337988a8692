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


@pytest.mark.parametrize("decoder", ["gather", "gather_bulk", "tc"])
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
    # "gather": rows through registers (default); "gather_bulk": through bulk-copy shared-memory rings (csrc/bulk.cuh)
    org.set_gather_mode("bulk" if decoder == "gather_bulk" else "ldg")
    assert org.gather_mode() == ("bulk" if decoder == "gather_bulk" else "ldg")
    decoder = "gather" if decoder == "gather_bulk" else decoder
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


def zipf_csr(rng, n_rows, n_cols, per_row_lo, per_row_hi, heavy_rows=(), heavy_len=1500, values="normal"):
    """Benchmark-like sparsity: row lengths spread between per_row_lo and per_row_hi (plus a few heavy rows, the
    ML1M users with ~2000 ratings), columns drawn without replacement from a Zipf popularity law so that the popular
    columns appear in far more than 64 rows of one 500-row batch."""
    w = 1.0 / np.arange(1, n_cols + 1) ** 0.9
    w = w[rng.permutation(n_cols)]
    w /= w.sum()
    lens = rng.integers(per_row_lo, per_row_hi + 1, size=n_rows)
    for r in heavy_rows:
        lens[r] = heavy_len
    indptr = np.zeros(n_rows + 1, np.int64)
    indptr[1:] = np.cumsum(lens)
    indices = np.empty(indptr[-1], np.int32)
    for r in range(n_rows):
        indices[indptr[r]:indptr[r + 1]] = np.sort(rng.choice(n_cols, size=lens[r], replace=False, p=w))
    if values == "rating":
        data = rng.integers(1, 6, size=indptr[-1]).astype(np.float32)
    else:
        data = rng.normal(size=indptr[-1]).astype(np.float32)
        data[data == 0] = 0.5
    return csr_matrix((data, indices, indptr.astype(np.int32)), shape=(n_rows, n_cols))


def chunk_census(T, batches, dec_chunk=128, seg_chunk=64):
    """(rows that span several decoder chunks, (batch, column) segments that span several reduction chunks)."""
    multi_rows = multi_segs = 0
    for rows in batches:
        rows = np.asarray(rows)
        tl = T.indptr[rows + 1] - T.indptr[rows]
        multi_rows += int((tl > dec_chunk).sum())
        cols = np.concatenate([T.indices[T.indptr[r]:T.indptr[r + 1]] for r in rows])
        multi_segs += int((np.bincount(cols) > seg_chunk).sum())
    return multi_rows, multi_segs


@pytest.mark.parametrize("mode", ["epoch", "epoch_fanout", "round", "round_pdl"])
@pytest.mark.parametrize("decoder", ["gather", "gather_bulk", "gather_classic", "tc"])
@pytest.mark.parametrize("shape", ["ml1m_batch", "zipf_wide"])
def test_train_benchmark_shape(nat, shape, decoder, mode):
    """The paths the benchmark runs (VERDICT r1 'parity hole'): 500-row batches over 3706 target columns with rows of
    far more than 128 targets (several decoder chunks + ae_decoder_finish summing dz_part) and popular columns in far
    more than 64 rows of a batch (multi-chunk segments + segment_finish), against the oracle's Organization.train /
    predict (reference src/organization.py:140-217, src/models/ae.py:135-142). Per-epoch plans with and without the
    backward fan-out, and the whole-round plan (one plan, one graph for all epochs)."""
    rng = np.random.default_rng(11 if shape == "ml1m_batch" else 12)
    if shape == "ml1m_batch":   # two ML1M-shape batches per epoch: 1000 x 3706, ~150 targets per row
        n_rows, n_enc, n_dec, bs, lo, hi = 1000, 206, 3706, 500, 20, 280
        heavy = (3, 500, 777)
    else:                       # fewer, much longer rows over a wider item space; three batches, the last one ragged
        n_rows, n_enc, n_dec, bs, lo, hi = 700, 160, 6000, 300, 150, 700
        heavy = (5, 650)
    D = rand_csr(rng, n_rows, n_enc, 0.03, empty_rows=(7, 8))
    T = zipf_csr(rng, n_rows, n_dec, lo, hi, heavy_rows=heavy, heavy_len=1900)
    torch.manual_seed(5)
    p0 = replay.init_ae_params(n_enc, n_dec)
    n_epochs = 2
    epoch_batches, masks, keep_epochs = [], [], []
    for e in range(n_epochs):
        perm = rng.permutation(n_rows)
        batches = [np.sort(perm[s:s + bs]) for s in range(0, n_rows, bs)]
        epoch_batches.append(batches)
        ke = [torch.from_numpy((rng.random((len(b), 128)) < 0.5).astype(np.float32)) for b in batches]
        keep_epochs.append(ke)
        masks += ke
    multi_rows, multi_segs = chunk_census(T, epoch_batches[0])
    assert multi_rows > 100 and multi_segs > 100, (multi_rows, multi_segs)  # the multi-chunk paths really run
    ref_p, ref_losses = train.train_org_ae(p0, D, T, "user", "explicit", epoch_batches, masks)
    d_csr = (cu(D.indptr, torch.int32), cu(D.indices, torch.int32), cu(D.data))
    t_csr = (cu(T.indptr, torch.int32), cu(T.indices, torch.int32))
    # "round_pdl": whole-round plan with the step's kernels launched with programmatic stream serialization
    org = nat.Org(n_rows, n_enc, n_dec, 256, 128, d_csr, t_csr, bs, 0,
                  plan_epochs=n_epochs if mode.startswith("round") else 1)
    org.set_pdl(mode == "round_pdl")
    # "gather": the fused six-launch step (csrc/fused.cu, the default); "gather_bulk": the same step with the gathered
    # rows travelling through bulk-copy rings (csrc/bulk.cuh); "gather_classic": one kernel per layer
    want_fused = decoder in ("gather", "gather_bulk")
    org.set_gather_mode("bulk" if decoder == "gather_bulk" else "ldg")
    if decoder == "gather_bulk":
        decoder = "gather"
    if decoder == "gather_classic":
        org.set_step_mode("classic")
        decoder = "gather"
    org.set_decoder_mode(decoder)
    assert org.step_mode() == ("fused" if want_fused else "classic")
    org.set_fanout(mode == "epoch_fanout")
    org.wait_current()
    org.set_params(flat_params(p0).cuda())
    tval = cu(T.data)
    org.set_target(tval)
    nb = len(epoch_batches[0])
    if mode.startswith("round"):
        rows_a = np.concatenate([np.concatenate(b) for b in epoch_batches])
        off = np.concatenate([[0], np.cumsum([len(r) for b in epoch_batches for r in b])])
        keep = torch.cat([k for ke in keep_epochs for k in ke]).to(torch.uint8).cuda()
        loss = torch.zeros(nb * n_epochs, device="cuda")
        org.train_epoch(cu(rows_a, torch.int32), cu(off, torch.int32), int(T.nnz) * n_epochs, int(D.nnz) * n_epochs,
                        keep=keep, epoch_loss=loss)
        org.sync()
        got_losses = loss.cpu().tolist()
    else:
        got_losses = []
        for e, batches in enumerate(epoch_batches):
            rows_a = np.concatenate(batches)
            off = np.concatenate([[0], np.cumsum([len(r) for r in batches])])
            keep = torch.cat(keep_epochs[e]).to(torch.uint8).cuda()
            loss = torch.zeros(nb, device="cuda")
            org.train_epoch(cu(rows_a, torch.int32), cu(off, torch.int32), int(T.nnz), int(D.nnz), keep=keep,
                            epoch_loss=loss)
            org.sync()
            got_losses += loss.cpu().tolist()
    assert len(got_losses) == len(ref_losses)
    assert rel_err(got_losses, ref_losses) < 1e-5
    got_p = unflat_params(org.get_params().cpu(), n_enc, n_dec)
    org.sync()
    tol_p = 5e-4 if decoder == "gather" else 1e-3
    for k, v in ref_p.items():
        assert rel_err(got_p[k], v) < tol_p, k
    T2 = zipf_csr(rng, n_rows, n_dec, 5, 60, heavy_rows=(1,), heavy_len=900)
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
