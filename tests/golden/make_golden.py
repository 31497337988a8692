"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py            # all cases, one subprocess each
    python tests/golden/make_golden.py --case X   # one case in this process

The reference (/root/reference/src, pure Python on torch CPU) cannot travel to the GPU box, so its
inputs/outputs on small seeded synthetic datasets are frozen here as .npz files. tests/ compare the
oracle (oracle/, CPU restatement) and the CUDA path against them. Reference call sites are cited at
each capture point. Nothing here is imported by the product, the GPU tests, smoke() or bench.py.
"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

SCRATCH = "/tmp/dmt_golden"

# case name -> (kind, control_name, synthetic dataset name)
CASES = {
    # model-level: forward / loss / grads / 3 Adam+clip steps on reference-built batches
    "model_mf_user_explicit": ("model", "ML100K_user_explicit_mf_0_genre_joint", "tiny-ML100K"),
    "model_mf_user_implicit_info1": ("model", "ML100K_user_implicit_mf_1_genre_joint", "tiny-ML100K"),
    "model_mlp_user_explicit": ("model", "ML100K_user_explicit_mlp_0_genre_joint", "tiny-ML100K"),
    "model_nmf_item_implicit": ("model", "ML100K_item_implicit_nmf_0_random-8_alone", "tiny-ML100K"),
    "model_nmf_item_implicit_info1_amazon": ("model", "Amazon_item_implicit_nmf_1_random-8_alone", "tiny-Amazon"),
    "model_ae_user_explicit": ("model", "ML100K_user_explicit_ae_0_genre_assist_constant-0.1_constant", "tiny-ML100K"),
    "model_ae_user_implicit_douban": ("model", "Douban_user_implicit_ae_0_genre_assist_constant-0.3_constant",
                                      "tiny-Douban"),
    "model_ae_item_explicit": ("model", "ML100K_item_explicit_ae_0_random-4_assist_constant-0.1_constant",
                               "tiny-ML100K"),
    # MTAL-level: make_dataset residuals and update() under every ar/aw/match_rate/pl variant
    "mtal_douban_user_explicit": ("mtal", "Douban_user_explicit_ae_0_genre_assist_constant-0.3_constant",
                                  "tiny-Douban"),
    "mtal_amazon_user_implicit_dp": ("mtal", "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_optim_0.5_dp-10",
                                     "tiny-Amazon"),
    "mtal_ml_user_explicit": ("mtal", "ML100K_user_explicit_ae_0_genre_assist_constant-0.1_constant", "tiny-ML100K"),
    # round-level: a whole (shortened) experiment, per-round global outputs and metrics
    "round_douban_user_explicit": ("round", "Douban_user_explicit_ae_0_genre_assist_constant-0.3_constant",
                                   "tiny-Douban"),
    "round_amazon_user_implicit": ("round", "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_optim_0.5_dp-10",
                                   "tiny-Amazon"),
    "round_ml_item_explicit": ("round", "ML100K_item_explicit_ae_0_random-4_assist_optim-0.1_constant", "tiny-ML100K"),
    # cold start (12-field control name, src/train_recsys_assist.py:52-56,180-182; src/assist.py:109-117,150-157;
    # src/models/assist.py:28-34): organization 0 holds only the first half of the aligned rows
    "round_ml_user_explicit_cs": ("round", "ML100K_user_explicit_ae_0_genre_assist_constant-0.3_constant_1_none_0.5",
                                  "tiny-ML100K"),
    # joint driver (src/train_recsys_joint.py:40-199): joint training epochs, models.distribute, the per-organization
    # test() loop with its combined metrics
    "joint_mf_user_explicit": ("joint", "ML100K_user_explicit_mf_0_genre_joint", "tiny-ML100K"),
    "joint_nmf_user_implicit": ("joint", "ML100K_user_implicit_nmf_0_genre_joint", "tiny-ML100K"),
}


def put_csr(out, name, m):
    m = m.tocsr()
    out[name + "/indptr"] = m.indptr.astype(np.int64)
    out[name + "/indices"] = m.indices.astype(np.int64)
    out[name + "/data"] = m.data.astype(np.float32)
    out[name + "/shape"] = np.array(m.shape, dtype=np.int64)


def put_dict(out, prefix, d):
    import torch

    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out["{}/{}".format(prefix, k)] = v.detach().cpu().numpy().copy()
        elif isinstance(v, (bool, int, float)):
            out["{}/{}".format(prefix, k)] = np.array(v)


def setup(case):
    import torch
    from dmtcdr_b200 import synth
    import ref_harness

    kind, control, data_name = CASES[case]
    work = os.path.join(SCRATCH, case)
    data = synth.make_rating_data(data_name, seed=0)
    synth.write_reference_layout(data, os.path.join(work, "data"))
    run_mode = control.split("_")[6]
    driver = {"assist": "train_recsys_assist", "joint": "train_recsys_joint", "alone": "train_recsys_alone"}[run_mode]
    ns = ref_harness.enter(work, control, seed=0, driver_name=driver)
    torch.manual_seed(0)
    return kind, control, ns


def start_assist(ns, local_epochs, rounds):
    """Reference flow up to round 0 (src/train_recsys_assist.py:44-80)."""
    cfg = ns.cfg
    cfg["local"]["num_epochs"] = local_epochs
    cfg["global"]["num_epochs"] = rounds
    dataset = ns.data.fetch_dataset(cfg["data_name"], verbose=False)
    ns.utils.process_dataset(dataset)
    data_split = ns.data.split_dataset(dataset)
    dataset = ns.data.make_split_dataset(data_split)
    if "cs" in cfg:  # the driver's own cold-start truncation (src/train_recsys_assist.py:52-56)
        start_size = int(len(dataset[0]["train"]) * cfg["cs"])
        dataset[0]["train"].data = dataset[0]["train"].data[:start_size]
        dataset[0]["train"].target = dataset[0]["train"].target[:start_size]
    assist = ns.assist.Assist(data_split)
    organization = assist.make_organization()
    names = ["Loss", "RMSE"] if cfg["target_mode"] == "explicit" else ["Loss", "NDCG"]
    metric = ns.metrics.Metric({"train": names, "test": names})
    logger = ns.logger.make_logger("output/runs/golden")
    ns.driver.initialize(dataset, assist, organization, metric, logger, 0)
    ns.driver.test(assist, metric, logger, 0)
    m0 = {k: float(v) for k, v in logger.mean.items() if k.startswith("test/")}
    logger.reset()
    return dataset, data_split, assist, organization, metric, logger, m0


def meta_of(ns, extra=None):
    cfg = ns.cfg
    keys = ["data_name", "data_mode", "target_mode", "model_name", "info", "num_organizations", "control_name"]
    m = {k: cfg[k] for k in keys if k in cfg}
    m["info_size"] = cfg.get("info_size")
    m["assist"] = {k: v for k, v in cfg["assist"].items() if k in ("ar", "ar_mode", "aw_mode", "match_rate")}
    for k in ("pl", "pl_mode", "pl_param"):
        if k in cfg:
            m[k] = cfg[k]
    m["batch_size"] = cfg[cfg["model_name"]]["batch_size"]["train"]
    m["num_users"] = cfg.get("num_users")
    m["num_items"] = cfg.get("num_items")
    if extra:
        m.update(extra)
    return m


def case_model(case, ns):
    """Reference model forward/backward on reference-built batches.
    mf/mlp/nmf: src/models/{mf,mlp,nmf}.py forward; batches from src/data.py PairInput + DataLoader.
    ae: src/models/ae.py:98-157 on FlatInput batches of one organization after make_dataset (src/assist.py:43-79)."""
    import torch

    cfg = ns.cfg
    out = {}
    model_name = cfg["model_name"]
    extra = {}
    if model_name == "ae":
        dataset, data_split, assist, organization, metric, logger, _ = start_assist(ns, 1, 1)
        dataset = assist.make_dataset(dataset, 1)
        i = 1
        ds = dataset[i]["train"]
        loader = ns.data.make_data_loader({"train": ds}, "local")["train"]
        model = ns.models.ae(ds.num_users["data"], ds.num_items["data"], ds.num_users["target"],
                             ds.num_items["target"])
        extra = {"enc_users": ds.num_users["data"], "enc_items": ds.num_items["data"],
                 "dec_users": ds.num_users["target"], "dec_items": ds.num_items["target"]}
        opt_tag = "local"
    else:
        dataset = ns.data.fetch_dataset(cfg["data_name"], verbose=False)
        ns.utils.process_dataset(dataset)
        data_split = ns.data.split_dataset(dataset)
        if cfg["run_mode"] == "joint":
            loader = ns.data.make_data_loader(dataset, model_name)["train"]
            model = getattr(ns.models, model_name)()
            extra = {"n_users": cfg["num_users"]["data"], "n_items": cfg["num_items"]["data"]}
        else:
            local = ns.data.make_split_dataset(data_split)
            ds = local[0]
            loader = ns.data.make_data_loader(ds, model_name)["train"]
            nu, ni = ds["train"].num_users["data"], ds["train"].num_items["data"]
            model = getattr(ns.models, model_name)(nu, ni)
            extra = {"n_users": nu, "n_items": ni}
        opt_tag = model_name
    put_dict(out, "sd0", model.state_dict())
    batches = []
    for b in loader:
        b = ns.utils.collate(b)
        if len(b[cfg["data_mode"]]) == 0:
            continue
        batches.append(b)
    batches = [batches[0], batches[-1]]
    masks = []
    for j, b in enumerate(batches):
        put_dict(out, "b{}/in".format(j), b)
        model.train(True)
        model.zero_grad()
        if model_name == "ae":
            b["local"] = True
        st = torch.get_rng_state()
        o = model(b)
        o["loss"].backward()
        if model_name == "ae":
            # nn.Dropout(0.5) on CPU == x * bernoulli_(0.5) / 0.5 drawn from the global generator
            after = torch.get_rng_state()
            torch.set_rng_state(st)
            nrow = len(torch.unique(torch.cat([b[cfg["data_mode"]], b["target_" + cfg["data_mode"]]])))
            mask = torch.empty(nrow, cfg["ae"]["encoder_hidden_size"][-1]).bernoulli_(0.5)
            torch.set_rng_state(after)
            out["b{}/mask".format(j)] = mask.numpy().astype(np.uint8)
            masks.append(mask)
        out["b{}/train/target_rating".format(j)] = o["target_rating"].detach().numpy()
        out["b{}/train/loss".format(j)] = o["loss"].detach().numpy()
        for n, p in model.named_parameters():
            out["b{}/grad/{}".format(j, n)] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        model.train(False)
        with torch.no_grad():
            if model_name == "ae":
                b["local"] = False
            o = model(b)
        out["b{}/eval/target_rating".format(j)] = o["target_rating"].numpy()
        out["b{}/eval/loss".format(j)] = o["loss"].numpy()
    # 4 optimizer steps: clip_grad_norm_(.,1) + Adam(lr 1e-3, wd 5e-4) (src/organization.py:158-162, src/utils.py:253-254)
    model.train(True)
    optimizer = ns.utils.make_optimizer(model, opt_tag)
    losses = []
    for step in range(4):
        b = batches[step % 2]
        if model_name == "ae":
            b["local"] = True
            # replay the recorded mask: re-seed so that dropout draws exactly masks[step % 2]
            torch.manual_seed(1000 + step)
            st = torch.get_rng_state()
            nrow = masks[step % 2].shape[0]
            m = torch.empty(nrow, cfg["ae"]["encoder_hidden_size"][-1]).bernoulli_(0.5)
            out["steps/mask{}".format(step)] = m.numpy().astype(np.uint8)
            torch.set_rng_state(st)
        optimizer.zero_grad()
        o = model(b)
        o["loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optimizer.step()
        losses.append(float(o["loss"]))
    out["steps/loss"] = np.array(losses, dtype=np.float32)
    put_dict(out, "sd4", model.state_dict())
    out["meta"] = np.array(json.dumps(meta_of(ns, extra)))
    return out


def capture_round_inputs(out, ns, dataset, assist, organization, metric, logger, t):
    """One reference round up to (not including) update(); src/train_recsys_assist.py:82-84."""
    dataset = assist.make_dataset(dataset, t)
    for k in ("train", "test"):
        out["r{}/residual/{}".format(t, k)] = dataset[0][k].target.data.astype(np.float32)
    ns.driver.train(dataset, organization, metric, logger, t)
    outs = ns.driver.gather(dataset, organization, t)
    for k in ("train", "test"):
        ref = assist.organization_target[0][k]
        for j, o in enumerate(outs):
            vals = o[k].data.astype(np.float32)
            if o[k].nnz != ref.nnz:
                # cold start: the organization predicted only the rows it holds; the rest of the global pattern is
                # absent from its output (the reference pads it with NaN at src/assist.py:109-111)
                assert np.array_equal(o[k].indices, ref.indices[:o[k].nnz])
                vals = np.concatenate([vals, np.full(ref.nnz - o[k].nnz, np.nan, np.float32)])
            else:
                assert np.array_equal(o[k].indptr, ref.indptr) and np.array_equal(o[k].indices, ref.indices)
            out["r{}/org_out/{}/{}".format(t, k, j)] = vals
    return dataset, outs


def case_mtal(case, ns):
    """make_dataset (src/assist.py:43-79) and update (src/assist.py:81-179) under every ar/aw/match_rate variant,
    all on the same captured organization outputs."""
    import copy
    import torch

    cfg = ns.cfg
    out = {}
    dataset, data_split, assist, organization, metric, logger, _ = start_assist(ns, 1, 1)
    for i, s in enumerate(data_split):
        out["data_split/{}".format(i)] = s.numpy().astype(np.int64)
    for k in ("train", "test"):
        put_csr(out, "F0/" + k, assist.organization_output[0][k])
        put_csr(out, "y/" + k, assist.organization_target[0][k])
    # residual without privacy noise, for the pure elementwise check
    pl = cfg.pop("pl", None)
    ds_nopl = assist.make_dataset(copy.deepcopy(dataset), 1)
    for k in ("train", "test"):
        out["residual_nopl/{}".format(k)] = ds_nopl[0][k].target.data.astype(np.float32)
    if pl is not None:
        cfg["pl"] = pl
    dataset, outs = capture_round_inputs(out, ns, dataset, assist, organization, metric, logger, 1)
    base_assist = dict(cfg["assist"])
    variants = {
        "const_const": {"ar_mode": "constant", "aw_mode": "constant", "match_rate": 1.0},
        "optim_const": {"ar_mode": "optim", "aw_mode": "constant", "match_rate": 1.0},
        "const_optim": {"ar_mode": "constant", "aw_mode": "optim", "match_rate": 1.0},
        "optim_optim": {"ar_mode": "optim", "aw_mode": "optim", "match_rate": 1.0},
        "const_optim_match0.5": {"ar_mode": "constant", "aw_mode": "optim", "match_rate": 0.5},
        "const_const_match0.5": {"ar_mode": "constant", "aw_mode": "constant", "match_rate": 0.5},
    }
    for name, v in variants.items():
        cfg["assist"].update(v)
        torch.manual_seed(7)
        assist.update(outs, 1)
        for k in ("train", "test"):
            m = assist.organization_output[1][k]
            ref = assist.organization_target[0][k]
            assert np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices)
            out["update/{}/F1/{}".format(name, k)] = m.data.astype(np.float32)
        for i in range(len(data_split)):
            sd = assist.ar_state_dict[1][i]
            out["update/{}/rate/{}".format(name, i)] = sd["assist_rate"].numpy()
            out["update/{}/weight/{}".format(name, i)] = sd["assist_weight"].numpy()
    cfg["assist"].clear()
    cfg["assist"].update(base_assist)
    out["variants"] = np.array(json.dumps(variants))
    out["meta"] = np.array(json.dumps(meta_of(ns)))
    return out


def case_round(case, ns):
    """Whole shortened experiment: 2 rounds x 2 local epochs (src/train_recsys_assist.py:78-93)."""
    cfg = ns.cfg
    out = {}
    dataset, data_split, assist, organization, metric, logger, m0 = start_assist(ns, 2, 2)
    for i, s in enumerate(data_split):
        out["data_split/{}".format(i)] = s.numpy().astype(np.int64)
    metrics = {0: m0}
    for k in ("train", "test"):
        put_csr(out, "F0/" + k, assist.organization_output[0][k])
        put_csr(out, "y/" + k, assist.organization_target[0][k])
    for t in (1, 2):
        dataset, outs = capture_round_inputs(out, ns, dataset, assist, organization, metric, logger, t)
        assist.update(outs, t)
        ns.driver.test(assist, metric, logger, t)
        metrics[t] = {k: float(v) for k, v in logger.mean.items() if k.startswith("test/")}
        logger.reset()
        for k in ("train", "test"):
            out["F{}/{}".format(t, k)] = assist.organization_output[t][k].data.astype(np.float32)
        for i in range(len(data_split)):
            sd = assist.ar_state_dict[t][i]
            out["ar{}/rate/{}".format(t, i)] = sd["assist_rate"].numpy()
            out["ar{}/weight/{}".format(t, i)] = sd["assist_weight"].numpy()
    put_dict(out, "org0_sd1", organization[0].model_state_dict[1])
    put_dict(out, "org0_base", organization[0].model_state_dict[0])
    out["metrics"] = np.array(json.dumps(metrics))
    out["meta"] = np.array(json.dumps(meta_of(ns, {"local_epochs": 2, "rounds": 2})))
    return out


def case_joint(case, ns):
    """Two epochs of the joint driver's own loop (src/train_recsys_joint.py:93-97: train, models.distribute, test) on
    the tiny dataset: the initial joint state_dict, every epoch's sampler order, the state_dict after every epoch and
    the test metrics the driver logs from the DISTRIBUTED local models (per-organization Loss + combined RMSE/NDCG)."""
    import copy
    import torch

    cfg = ns.cfg
    out = {}
    name = cfg["model_name"]
    dataset = ns.data.fetch_dataset(cfg["data_name"], verbose=False)
    ns.utils.process_dataset(dataset)
    data_split = ns.data.split_dataset(dataset)
    for i, s in enumerate(data_split):
        out["data_split/{}".format(i)] = s.numpy().astype(np.int64)
    data_loader = ns.data.make_data_loader(dataset, name)
    model = eval("ns.models.{}()".format(name))
    put_dict(out, "sd0", model.state_dict())
    optimizer = ns.utils.make_optimizer(model, name)
    names = ["Loss", "RMSE"] if cfg["target_mode"] == "explicit" else ["Loss", "NDCG"]
    metric = ns.metrics.Metric({"train": names, "test": names})
    logger = ns.logger.make_logger("output/runs/golden")
    local_dataset = ns.data.make_split_dataset(data_split)
    local_loader, local_model = [], []
    for i in range(len(local_dataset)):
        local_loader.append(ns.data.make_data_loader(local_dataset[i], name)["test"])
        nu = local_dataset[i]["train"].num_users["data"]
        ni = local_dataset[i]["train"].num_items["data"]
        local_model.append(eval("ns.models.{}(nu, ni)".format(name)))
    metrics = {}
    for epoch in (1, 2):
        st = torch.get_rng_state()
        it = iter(data_loader["train"])
        order = [np.asarray(b, dtype=np.int64) for b in it._sampler_iter]
        torch.set_rng_state(st)
        out["e{}/rows".format(epoch)] = np.concatenate(order)
        out["e{}/batch_sizes".format(epoch)] = np.array([len(b) for b in order], dtype=np.int64)
        ns.driver.train(data_loader["train"], model, optimizer, metric, logger, epoch)
        ns.models.distribute(model, local_model, data_split)
        ns.driver.test(local_loader, data_split, local_model, metric, logger, epoch)
        metrics[epoch] = {k: float(v) for k, v in logger.mean.items()}
        logger.reset()
        put_dict(out, "sd{}".format(epoch), model.state_dict())
    put_dict(out, "local0_sd2", local_model[0].state_dict())
    out["metrics"] = np.array(json.dumps(metrics))
    out["meta"] = np.array(json.dumps(meta_of(ns, {"epochs": 2})))
    return out


def run_case(case):
    kind, control, ns = setup(case)
    fn = {"model": case_model, "mtal": case_mtal, "round": case_round, "joint": case_joint}[kind]
    out = fn(case, ns)
    path = os.path.join(HERE, case + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", len(out), "arrays")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    a = ap.parse_args()
    if a.case:
        run_case(a.case)
    else:
        for c in CASES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], capture_output=True, text=True)
            tail = r.stdout.strip().splitlines()[-1:] if r.returncode == 0 else [r.stderr[-2000:]]
            print(c, "rc=", r.returncode, *tail)
