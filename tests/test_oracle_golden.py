"""Pin the CPU oracle (oracle/) against the golden vectors produced by the unmodified reference.

Tolerances: loss <= 1e-5 relative, metrics <= 1e-4 absolute (BASELINE.json north_star); vectors of predictions,
gradients and parameters <= 2e-5 relative to the largest reference magnitude (fp32, different summation order).
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import models as om
from oracle import mtal, replay, train
from golden_io import Fixture, cases, rel_err

TOL_VEC = 2e-5
TOL_LOSS = 1e-5
TOL_ADAM = 5e-4  # parameters after optimizer steps, relative to max |w|
TOL_FIT = 2e-3  # L-BFGS-fitted rate / weight (its 1e-9 loss-change stopping rule is rounding-sensitive)


def batch_of(fx, j):
    return {k: v for k, v in fx.group("b{}/in".format(j)).items()}


def forward(fx, p, b, training, keep=None):
    m = fx.meta
    if m["model_name"] == "ae":
        return om.ae_forward(p, b, m["data_mode"], m["target_mode"], training, keep, local=training)
    return om.pair_forward(m["model_name"], p, b, m["target_mode"], training)


@pytest.mark.parametrize("case", cases("model"))
def test_model_forward_backward(case):
    fx = Fixture(case)
    sd = fx.group("sd0")
    for j in (0, 1):
        b = batch_of(fx, j)
        keep = torch.from_numpy(fx["b{}/mask".format(j)].astype(np.float32)) if "b{}/mask".format(j) in fx else None
        p = train.leaf(sd)
        pred, loss = forward(fx, p, b, True, keep)
        assert rel_err(pred.detach(), fx["b{}/train/target_rating".format(j)]) < TOL_VEC
        assert abs(float(loss) - float(fx["b{}/train/loss".format(j)])) <= TOL_LOSS * abs(float(loss))
        g = train.grads_of(loss, p)
        for name, ref in fx.group("b{}/grad".format(j), as_torch=False).items():
            got = np.zeros_like(ref) if g[name] is None else g[name].numpy()
            assert rel_err(got, ref) < 5e-5 or np.abs(got - ref).max() < 1e-9, name
        with torch.no_grad():
            pred, loss = forward(fx, sd, b, False)
        assert rel_err(pred, fx["b{}/eval/target_rating".format(j)]) < TOL_VEC
        assert abs(float(loss) - float(fx["b{}/eval/loss".format(j)])) <= TOL_LOSS * abs(float(loss))


@pytest.mark.parametrize("case", cases("model"))
def test_model_optimizer_steps(case):
    """4 steps of backward + clip_grad_norm_(1) + Adam(lr 1e-3, wd 5e-4)."""
    fx = Fixture(case)
    opt = train.Adam({k: v.clone() for k, v in fx.group("sd0").items()})
    losses = []
    for step in range(4):
        b = batch_of(fx, step % 2)
        keep = None
        if "steps/mask{}".format(step) in fx:
            keep = torch.from_numpy(fx["steps/mask{}".format(step)].astype(np.float32))
        losses.append(train.train_step(opt, lambda p: forward(fx, p, b, True, keep)))
    assert rel_err(losses, fx["steps/loss"]) < TOL_LOSS
    for name, ref in fx.group("sd4", as_torch=False).items():
        # Adam's m/sqrt(v) amplifies rounding noise of near-zero gradients: looser than TOL_VEC
        assert rel_err(opt.p[name].numpy(), ref) < TOL_ADAM, name


@pytest.mark.parametrize("case", cases("mtal"))
def test_residual_and_update(case):
    fx = Fixture(case)
    m = fx.meta
    clamp = mtal.needs_clamp(m["data_name"], m["data_mode"], m["target_mode"])
    y = {k: fx.csr("y/" + k) for k in ("train", "test")}
    F0 = {k: fx.csr("F0/" + k).data for k in y}
    for k in y:
        assert np.array_equal(fx.csr("F0/" + k).indices, y[k].indices)
        r = mtal.residual(F0[k], y[k].data, m["target_mode"], clamp)
        assert rel_err(r, fx["residual_nopl/" + k]) < 1e-6
    if m.get("pl", "none") != "none":
        np.random.seed(0)  # make_data_loader re-seeds numpy before every make_dataset (reference src/data.py:76)
        for k in ("train", "test"):
            r = mtal.residual(F0[k], y[k].data, m["target_mode"], clamp)
            r = mtal.dp(r, m["pl_param"]).astype(np.float32)
            assert rel_err(r, fx["r1/residual/" + k]) < 1e-6
    K = m["num_organizations"]
    split = [fx["data_split/{}".format(i)] for i in range(K)]
    org_out = [{k: fx["r1/org_out/{}/{}".format(k, j)] for k in y} for j in range(K)]
    indices = {k: y[k].indices for k in y}
    for name, v in fx.json("variants").items():
        Fn, fitted = mtal.update(F0, {k: y[k].data for k in y}, org_out, indices, split, y["train"].shape[1],
                                 m["target_mode"], m["assist"]["ar"], v["ar_mode"], v["aw_mode"], v["match_rate"])
        for i in range(K):
            assert rel_err(fitted[i][0], fx["update/{}/rate/{}".format(name, i)]) < TOL_FIT, (name, i)
            assert rel_err(fitted[i][1], fx["update/{}/weight/{}".format(name, i)]) < TOL_FIT, (name, i)
        for k in y:
            assert rel_err(Fn[k], fx["update/{}/F1/{}".format(name, k)]) < TOL_VEC, (name, k)


@pytest.mark.parametrize("case", cases("round"))
def test_round_replay(case):
    """Whole shortened experiment, RNG-identical replay: per-round global outputs and test metrics."""
    from dmtcdr_b200 import synth

    fx = Fixture(case)
    m = fx.meta
    data = synth.make_rating_data("tiny-" + m["data_name"], seed=0)
    res = replay.run_experiment(data, m["control_name"], seed=0, local_epochs=m["local_epochs"], rounds=m["rounds"])
    K = m["num_organizations"]
    for i in range(K):
        assert np.array_equal(res["data_split"][i], fx["data_split/{}".format(i)])
    for k in ("train", "test"):
        assert np.array_equal(res["y"][k].indices, fx["y/{}/indices".format(k)])
        assert rel_err(res["F"][0][k], fx["F0/{}/data".format(k)]) < 1e-6
    for name, ref in fx.group("org0_sd1", as_torch=False).items():
        assert rel_err(res["org0_sd1"][name].numpy(), ref) < 1e-4, name
    for t in (1, 2):
        for k in ("train", "test"):
            assert rel_err(res["F"][t][k], fx["F{}/{}".format(t, k)]) < 5e-4, (t, k)
    gm = fx.json("metrics")
    for t in (0, 1, 2):
        for name, ref in gm[str(t)].items():
            assert abs(res["metrics"][t][name] - ref) <= 1e-4, (t, name)
