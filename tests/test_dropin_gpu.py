"""Drop-in API (models / Organization / Assist) on the GPU against the reference's golden vectors.

module level : same state_dict + same batch (+ same dropout draw) -> same target_rating, loss, .grad
round level  : same seed, RNG-identical replay ('dmt_rng' = 'reference') -> same organization_output[t] and the
               same test metrics after every round (RMSE / NDCG within 1e-4, BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

from golden_io import Fixture, cases, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200

    return dmtcdr_b200.use_dropin()


def set_cfg(fx):
    from dmtcdr_b200.config import make_cfg, cfg

    make_cfg(fx.meta["control_name"], device="cuda", seed=0)
    cfg["info_size"] = fx.meta["info_size"]
    cfg["num_users"], cfg["num_items"] = fx.meta["num_users"], fx.meta["num_items"]
    return cfg


MODULE_CASES = cases("model")


@pytest.mark.parametrize("case", MODULE_CASES)
def test_module_forward_backward(dropin, case):
    models, _, _ = dropin
    fx = Fixture(case)
    cfg = set_cfg(fx)
    m = fx.meta
    if m["model_name"] == "ae":
        model = models.ae(m["enc_users"], m["enc_items"], m["dec_users"], m["dec_items"])
    else:
        model = getattr(models, m["model_name"])(m["n_users"], m["n_items"])
    model.load_state_dict(fx.group("sd0"))
    model = model.cuda()
    for j in (0, 1):
        b = {k: v.cuda() for k, v in fx.group("b{}/in".format(j)).items() if k != "local"}
        model.train(True)
        model.zero_grad()
        if m["model_name"] == "ae":
            b["local"] = True
            model.keep_override = torch.from_numpy(fx["b{}/mask".format(j)])
        out = model(b)
        out["loss"].backward()
        assert rel_err(out["target_rating"].detach().cpu(), fx["b{}/train/target_rating".format(j)]) < 2e-5
        ref_loss = float(fx["b{}/train/loss".format(j)])
        assert abs(float(out["loss"]) - ref_loss) <= 1e-5 * abs(ref_loss)
        for name, p in model.named_parameters():
            ref = fx["b{}/grad/{}".format(j, name)]
            got = np.zeros_like(ref) if p.grad is None else p.grad.cpu().numpy()
            assert rel_err(got, ref) < 5e-5 or np.abs(got - ref).max() < 1e-9, name
        model.train(False)
        with torch.no_grad():
            if m["model_name"] == "ae":
                b["local"] = False
            out = model(b)
        assert rel_err(out["target_rating"].cpu(), fx["b{}/eval/target_rating".format(j)]) < 2e-5
        ref_loss = float(fx["b{}/eval/loss".format(j)])
        assert abs(float(out["loss"]) - ref_loss) <= 1e-5 * abs(ref_loss)


@pytest.mark.parametrize("case", MODULE_CASES)
def test_module_with_torch_optimizer(dropin, case):
    """The drivers own clip_grad_norm_ + torch.optim.Adam for these models (src/train_recsys_joint.py:129-134)."""
    models, _, _ = dropin
    fx = Fixture(case)
    set_cfg(fx)
    m = fx.meta
    if m["model_name"] == "ae":
        model = models.ae(m["enc_users"], m["enc_items"], m["dec_users"], m["dec_items"])
    else:
        model = getattr(models, m["model_name"])(m["n_users"], m["n_items"])
    model.load_state_dict(fx.group("sd0"))
    model = model.cuda()
    model.train(True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=5e-4)
    losses = []
    for step in range(4):
        b = {k: v.cuda() for k, v in fx.group("b{}/in".format(step % 2)).items() if k != "local"}
        if m["model_name"] == "ae":
            b["local"] = True
            model.keep_override = torch.from_numpy(fx["steps/mask{}".format(step)])
        opt.zero_grad()
        out = model(b)
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        opt.step()
        losses.append(float(out["loss"]))
    assert rel_err(losses, fx["steps/loss"]) < 1e-5
    for name, ref in fx.group("sd4", as_torch=False).items():
        assert rel_err(model.state_dict()[name].cpu().numpy(), ref) < 5e-4, name


@pytest.mark.parametrize("case", cases("round"))
def test_round_replay_vs_reference(dropin, case):
    from dmtcdr_b200 import runner, synth

    fx = Fixture(case)
    m = fx.meta
    data = synth.make_rating_data("tiny-" + m["data_name"], seed=0)
    res = runner.run_assist_experiment(data, m["control_name"], seed=0, local_epochs=m["local_epochs"],
                                       rounds=m["rounds"], rng="reference", keep_objects=True)
    K = m["num_organizations"]
    for i in range(K):
        assert np.array_equal(res["data_split"][i], fx["data_split/{}".format(i)])
    for k in ("train", "test"):
        assert np.array_equal(res["y"][k].indices, fx["y/{}/indices".format(k)])  # bit-exact structure
        assert rel_err(res["F"][0][k], fx["F0/{}/data".format(k)]) < 1e-6
    sd = res["organization"][0].model_state_dict[1]
    for name, ref in fx.group("org0_sd1", as_torch=False).items():
        assert rel_err(sd[name].numpy(), ref) < 5e-4, name
    for t in (1, 2):
        for k in ("train", "test"):
            assert rel_err(res["F"][t][k], fx["F{}/{}".format(t, k)]) < 5e-4, (t, k)
        for i in range(K):
            a = res["ar_state_dict"][t][i]
            assert rel_err(a["assist_rate"], fx["ar{}/rate/{}".format(t, i)]) < 2e-3
            assert rel_err(a["assist_weight"], fx["ar{}/weight/{}".format(t, i)]) < 2e-3
    gm = fx.json("metrics")
    for t in (0, 1, 2):
        for name, ref in gm[str(t)].items():
            assert abs(res["metrics"][t][name] - ref) <= 1e-4, (t, name, res["metrics"][t][name], ref)


@pytest.mark.parametrize("case", cases("joint"))
def test_joint_driver_loop_vs_reference(dropin, case):
    """The joint driver's epoch (src/train_recsys_joint.py:93-97: train -> models.distribute -> test) with the drop-in
    models on the GPU against what the reference's own loop produced (tests/golden/make_golden.py: case_joint): the
    sampler orders of the fixture, the driver-owned clip_grad_norm_ + torch.optim.Adam, `models.distribute` into the 18
    per-organization models, and the test() loop on the device (runner.joint_test_device). State after every epoch
    <= 5e-4, logged test Loss <= 1e-5 relative, RMSE / NDCG <= 1e-4."""
    models, _, _ = dropin
    from dmtcdr_b200 import runner, synth
    from dmtcdr_b200.config import cfg

    fx = Fixture(case)
    m = fx.meta
    set_cfg(fx)
    name = m["model_name"]
    data = synth.make_rating_data("tiny-" + m["data_name"], seed=0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    K = m["num_organizations"]
    data_split = [torch.from_numpy(fx["data_split/{}".format(i)]) for i in range(K)]
    local_dataset = runner.make_split_dataset(dataset, data_split)
    model = getattr(models, name)()
    model.load_state_dict(fx.group("sd0"))
    model = model.cuda()
    opt = torch.optim.Adam(model.parameters(), lr=cfg[name]["lr"], betas=cfg[name]["betas"],
                           weight_decay=cfg[name]["weight_decay"])
    local_model = []
    for i in range(K):
        ds = local_dataset[i]["train"]
        local_model.append(getattr(models, name)(ds.num_users["data"], ds.num_items["data"]).cuda())
    gm = fx.json("metrics")
    tr = dataset["train"]
    for epoch in (1, 2):
        rows = fx["e{}/rows".format(epoch)]
        model.train(True)
        s0 = 0
        for n in fx["e{}/batch_sizes".format(epoch)]:
            b = {k: v.cuda() for k, v in runner.pair_batch(tr, rows[s0:s0 + n]).items()}
            s0 += int(n)
            if len(b[cfg["data_mode"]]) == 0:
                continue
            opt.zero_grad()
            out = model(b)
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
            opt.step()
        for k, ref in fx.group("sd{}".format(epoch), as_torch=False).items():
            assert rel_err(model.state_dict()[k].cpu().numpy(), ref) < 5e-4, (epoch, k)
        models.distribute(model, local_model, data_split)
        got = runner.joint_test_device(local_model, local_dataset, data_split, dataset["test"].target)
        for key, val in got.items():
            ref = gm[str(epoch)][key]
            tol = 1e-5 * abs(ref) if key.endswith("Loss") else 1e-4
            assert abs(val - ref) <= tol, (epoch, key, val, ref)
    for k, ref in fx.group("local0_sd2", as_torch=False).items():
        assert rel_err(local_model[0].state_dict()[k].cpu().numpy(), ref) < 5e-4, k


def test_cpu_tensors_fail_loudly(dropin):
    """No CPU path: the models refuse host tensors instead of silently computing elsewhere."""
    models, _, _ = dropin
    from dmtcdr_b200 import native

    fx = Fixture("model_mf_user_explicit")
    set_cfg(fx)
    model = models.mf(fx.meta["n_users"], fx.meta["n_items"])
    b = dict(fx.group("b0/in"))
    with pytest.raises(native.NativeError):
        model(b)


def test_ml1m_shape_round_vs_oracle(dropin):
    """One assistance round at the BENCHMARK's own shape (synthetic ML1M: 6040 x 3706, 900 188 train ratings, 18 genre
    organizations, 500-row batches, 1 local epoch) through the drop-in API with the reference-identical RNG, against
    the oracle's replay of the same experiment (oracle/replay.py, pinned to the reference by tests/golden): global
    prediction F_1 within 5e-4 of its largest magnitude, test RMSE within 1e-4 (BASELINE.json north_star). Every batch
    here has rows with more than 128 targets and columns with more than 64 entries, i.e. the multi-chunk decoder /
    segment kernels that the timed rounds run (reference src/models/ae.py:135-142, src/organization.py:149-162)."""
    from dmtcdr_b200 import runner, synth
    from oracle import replay

    control = "ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant"
    data = synth.make_rating_data("ML1M", seed=0)
    got = runner.run_assist_experiment(data, control, seed=0, local_epochs=1, rounds=1, rng="reference")
    ref = replay.run_experiment(data, control, seed=0, local_epochs=1, rounds=1)
    for i in range(18):
        assert np.array_equal(got["data_split"][i], ref["data_split"][i])
    for t in (0, 1):
        for k in ("train", "test"):
            assert rel_err(got["F"][t][k], ref["F"][t][k]) < (1e-6 if t == 0 else 5e-4), (t, k)
        r_got, r_ref = got["metrics"][t]["test/RMSE"], ref["metrics"][t]["test/RMSE"]
        assert abs(r_got - r_ref) <= 1e-4, (t, r_got, r_ref)


@pytest.mark.parametrize("target_mode", ["explicit", "implicit"])
@pytest.mark.parametrize("cold", ["none", "tail", "scattered"])
@pytest.mark.parametrize("modes", [("optim", "optim"), ("optim", "constant"), ("constant", "optim")])
def test_models_assist_is_differentiable(dropin, modes, cold, target_mode):
    """models.assist as the reference uses it in its fit (closure at src/assist.py:121-126: forward, loss,
    loss.backward() through the module): target, loss and the .grad of assist_rate / assist_weight against the
    oracle's autograd form (oracle/models.py: assist_forward, reference src/models/assist.py:25-40), including the
    cold-start branch (NaN in slot 0 -> softmax(w[1:]), rows moved to the end)."""
    models, _, _ = dropin
    from dmtcdr_b200.config import cfg, make_cfg
    from oracle import models as om

    make_cfg("ML100K_user_{}_ae_0_genre_assist_{}-0.3_{}".format(target_mode, modes[0], modes[1]), device="cuda", seed=0)
    g = torch.Generator().manual_seed(3)
    n, K, n_rate = 6001, cfg["num_organizations"], 53
    out = torch.randn(n, K, generator=g)
    if cold == "tail":
        out[n * 2 // 3:, 0] = float("nan")
    elif cold == "scattered":
        out[torch.rand(n, generator=g) < 0.3, 0] = float("nan")
    h = torch.randn(n, generator=g)
    idx = torch.randint(0, n_rate - 3, (n,), generator=g)  # the last columns own no entry: their rate gradient is 0
    t = torch.randn(n, generator=g) if target_mode == "explicit" else (torch.rand(n, generator=g) < 0.4).float()
    rate0 = 0.3 + 0.1 * torch.randn(n_rate, generator=g)
    w0 = torch.randn(K, generator=g)
    module = models.assist(n_rate)
    with torch.no_grad():
        module.assist_rate.copy_(rate0)
        module.assist_weight.copy_(w0)
    module = module.to("cuda")
    module.train(True)
    inp = {"history": h.cuda(), "output": out.cuda(), "target": t.cuda(), "output_idx": idx.cuda()}
    res = module(inp)
    if any(p.requires_grad for p in module.parameters()):
        res["loss"].backward()
    rate = rate0.clone().requires_grad_(modes[0] == "optim")
    w = w0.clone().requires_grad_(modes[1] == "optim")
    pred, loss = om.assist_forward(rate, w, h, out, idx, t, target_mode)
    loss.backward()
    assert rel_err(res["target"].detach().cpu().numpy(), pred.detach().numpy()) < 2e-6
    assert abs(float(res["loss"]) - float(loss)) <= 1e-5 * abs(float(loss))
    if modes[0] == "optim":
        assert rel_err(module.assist_rate.grad.cpu().numpy(), rate.grad.numpy()) < 2e-5
    if modes[1] == "optim":
        assert rel_err(module.assist_weight.grad.cpu().numpy(), w.grad.numpy()) < 2e-5
    # a second evaluation of the same input dict (the L-BFGS closure) reuses the cached index work and agrees bit for bit
    res2 = module(inp)
    assert torch.equal(res2["target"], res["target"])
