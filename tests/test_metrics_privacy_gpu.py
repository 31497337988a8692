"""Device-side evaluation (dmt_eval_blocks) and privacy transforms (dmt_privacy) — SURVEY.md §8f rows 3 and 4.

Evaluation: against the oracle's restatement of the reference test loop (oracle/metrics.py: per-block Loss / RMSE /
dense NDCG@10, entry-weighted over blocks) on seeded ragged CSRs, and against the metrics the UNMODIFIED reference
logged for its own outputs (tests/golden/round_*.npz), within the north_star's 1e-4.
Privacy: the quantile pair bit-close to numpy's, the clip exact, the noise distribution by its moments and KS distance.
"""
import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from oracle import metrics as ometrics
from golden_io import Fixture, cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import engine, native

    native.load()
    return engine


def _state(E, y, target_mode):
    n_cols = y.shape[1]
    return E.MtalState({"test": y}, [np.arange(n_cols)], target_mode, "cuda")


@pytest.mark.parametrize("target_mode", ["explicit", "implicit"])
@pytest.mark.parametrize("n_rows,n_cols,density,block", [(130, 100, 0.08, 50), (700, 300, 0.05, 500), (64, 9, 0.5, 100),
                                                         (300, 2000, 0.2, 128)])
def test_eval_blocks_vs_oracle(E, target_mode, n_rows, n_cols, density, block):
    rng = np.random.default_rng(n_rows + n_cols)
    mask = rng.random((n_rows, n_cols)) < density
    mask[3] = False          # empty rows
    mask[n_rows // 2] = False
    if n_rows > 60:
        mask[50:60] = False  # (part of) a block without entries
    val = rng.integers(1, 6, size=(n_rows, n_cols)).astype(np.float32)
    if target_mode == "implicit":
        val = (val >= 3.5).astype(np.float32) + 2.0  # keep the structure, binarise below
    y = csr_matrix(val * mask)
    y.sort_indices()
    if target_mode == "implicit":
        y.data = y.data - 2.0
    F = rng.normal(size=y.nnz).astype(np.float32)
    ref = ometrics.evaluate(F, y, "user", target_mode, block)
    st = _state(E, y, target_mode)
    got = st.evaluate(torch.from_numpy(F).cuda(), "test", block)
    assert set(got) == set(ref)
    assert abs(got["test/Loss"] - ref["test/Loss"]) <= 1e-5 * abs(ref["test/Loss"])
    name = "test/RMSE" if target_mode == "explicit" else "test/NDCG"
    assert abs(got[name] - ref[name]) <= 1e-4, (got, ref)
    again = st.evaluate(torch.from_numpy(F).cuda(), "test", block)
    assert again == got  # no atomics, fixed orders


@pytest.mark.parametrize("case", cases("round"))
def test_eval_blocks_vs_reference_logged_metrics(E, case):
    """The reference's own logged test metrics for its own F_t (golden fixtures)."""
    fx = Fixture(case)
    y = fx.csr("y/test")
    mode = fx.meta["target_mode"]
    cold = len(fx.meta["control_name"].split("_")) >= 12  # cold-start run: scored on organization 0's columns only
    if cold:
        K = fx.meta["num_organizations"]
        st = E.MtalState({"test": y}, [fx["data_split/{}".format(i)] for i in range(K)], mode, "cuda")
    else:
        st = _state(E, y, mode)
    gm = fx.json("metrics")
    for t in range(fx.meta["rounds"] + 1):
        F = fx.csr("F0/test").data if t == 0 else fx["F{}/test".format(t)]
        F = np.asarray(F, dtype=np.float32)
        got = st.evaluate(torch.from_numpy(F).cuda(), "test", fx.meta["batch_size"], org=0 if cold else None)
        # Tied scores inside a row (round 0: the base predictor gives every rating of equally-rated items the same
        # value) make NDCG depend on torch.topk's unspecified tie order (it differs between torch's CPU and CUDA
        # kernels too); the kernel ranks ties by storage position. 1e-4 holds wherever the ranking is well defined.
        ties = any(len(np.unique(F[a:b])) < b - a for a, b in zip(y.indptr[:-1], y.indptr[1:]))
        for name, ref in gm[str(t)].items():
            if name in got:
                tol = 5e-3 if (ties and name.endswith("NDCG")) else 1e-4
                assert abs(got[name] - ref) <= tol, (t, name, got[name], ref)



@pytest.mark.parametrize("n", [1, 2, 1000, 200003])
def test_privacy_quantiles_clip_and_noise(E, n):
    from dmtcdr_b200 import native

    rng = np.random.default_rng(n)
    y = (rng.standard_t(3, size=n) * 0.7).astype(np.float32)
    a, b = np.quantile(y, 0.025), np.quantile(y, 0.975)
    yd = torch.from_numpy(y).cuda()
    out, q = native.privacy(yd, "dp", 10.0, seed=123)
    q = q.cpu().numpy()
    assert abs(q[0] - a) <= 1e-6 * max(1.0, abs(a)) and abs(q[1] - b) <= 1e-6 * max(1.0, abs(b))
    noise = out.cpu().numpy().astype(np.float64) - np.clip(y, q[0], q[1])
    scale = max(0.0, float(q[1] - q[0]) / 10.0)
    again, _ = native.privacy(yd, "dp", 10.0, seed=123)
    assert torch.equal(out, again)                      # reproducible for a seed
    other, _ = native.privacy(yd, "dp", 10.0, seed=124)
    if n >= 1000 and scale > 0:
        assert not torch.equal(out, other)
        # Laplace(0, s): mean 0, E|x| = s, var 2 s^2; Kolmogorov distance to the exact CDF
        assert abs(noise.mean()) < 5 * scale * np.sqrt(2.0 / n)
        assert abs(np.abs(noise).mean() / scale - 1.0) < 5.0 / np.sqrt(n)
        xs = np.sort(noise)
        cdf = np.where(xs < 0, 0.5 * np.exp(xs / scale), 1 - 0.5 * np.exp(-xs / scale))
        ks = np.abs(cdf - (np.arange(1, n + 1) - 0.5) / n).max()
        assert ks < 2.0 / np.sqrt(n)
    # interval privacy: unbiased around the clipped value's complement structure -> finite, inside [2a-b, 2b-a]
    ipo, q2 = native.privacy(yd, "ip", 3.0, seed=5)
    ipo = ipo.cpu().numpy()
    assert np.isfinite(ipo).all()
    lo, hi = 2 * q[0] - q[1], 2 * q[1] - q[0]
    assert (ipo >= lo - 1e-5 * max(1, abs(lo))).all() and (ipo <= hi + 1e-5 * max(1, abs(hi))).all()
    if n >= 1000:
        # E[ip(y)] for y inside [a, b] equals y (the estimator of src/privacy.py is unbiased there)
        inside = (y > q[0]) & (y < q[1])
        assert abs((ipo[inside] - y[inside]).mean()) < 6 * (q[1] - q[0]) / np.sqrt(inside.sum())


def test_sharded_rounds_with_device_privacy_and_eval(E):
    """AssistRounds with device-side dp noise: identical on emulated ranks (same seed -> same noise), metrics finite."""
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import roundloop, runner, synth
    from dmtcdr_b200.config import make_cfg

    control = "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_constant"
    make_cfg(control, device="cuda", seed=0)
    data = synth.make_rating_data("tiny-Amazon", seed=0)
    torch.manual_seed(0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    split = [s.numpy() for s in runner.split_dataset(dataset)]
    mats = {k: (dataset[k].data, dataset[k].target) for k in dataset}
    kw = dict(target_mode="implicit", batch_rows=50, clamp=True, ar=0.1, local_epochs=1, device="cuda:0", seed=3,
              privacy=("dp", 10.0))
    one = roundloop.AssistRounds(mats, split, rank=0, world=1, **kw)
    two = roundloop.AssistRounds(mats, split, rank=0, world=1, **kw)
    clean = None
    for r in (one, two):
        r.round0()
        clean = r.state.residual(r.F["train"], "train", True).clone()
        r.run_round(1)
        r.sync()
    for k in ("train", "test"):
        assert torch.equal(one.residual[k], two.residual[k])
        assert torch.equal(one.F[k], two.F[k])
    m = one.evaluate("test")
    assert np.isfinite(m["test/Loss"]) and 0.0 <= m["test/NDCG"] <= 1.0
    # the noise is really there: the broadcast residuals differ from the clean ones of F_0
    assert not torch.equal(clean, one.residual["train"])
    one.close()
    two.close()
