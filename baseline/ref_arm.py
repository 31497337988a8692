#!/usr/bin/env python
"""Reference arm: time the UNMODIFIED reference (baseline/_ref/src, a git-ignored copy made by
``__graft_entry__.build()`` from /root/reference/src; falls back to /root/reference/src in the build container)
through its own public API on the host CPU (or, informative, ``--device cuda`` = PyTorch eager on the B200).

    python baseline/ref_arm.py --control ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant \
        --data ML1M --device cpu --threads 16 --steps 3 --warmup 1

What runs inside the timed regions is reference code only — ``Assist.make_dataset`` (src/assist.py:43-79),
``Organization.train`` (src/organization.py:140-178), ``Organization.predict`` (:180-217), ``Assist.update``
(src/assist.py:81-179); none of this repo's models, kernels or engine is on that path (``import models`` resolves to
the reference's ``src/models``). Harness-side only (no reference file is edited):
  * stub ``matplotlib`` / ``anytree`` (imported by the reference, unused on the path),
  * scipy >= 1.8 fancy indexing with torch tensors, ``torch.load(weights_only=False)``,
  * the synthetic dataset written in the reference's pickle layout (synth.write_reference_layout),
  * round 0 (``initialize``: 154 s of per-user Python for 18 organizations, not part of the round metric) is replaced
    by filling ``assist.organization_output[0]`` / ``organization_target[0]`` with the global target CSR and the
    per-column mean, the same containers ``initialize`` fills (src/train_recsys_assist.py:98-141).

One timed "step" = ``Organization.train`` of ONE organization for ONE local epoch (cfg['local']['num_epochs'] = 1)
on the round's broadcast residuals; ``make_dataset``, ``predict`` (train+test) of one organization and ``update`` are
timed once each. The full-round figure composes them exactly as the reference's loop does
(src/train_recsys_assist.py:82-85,144-172):
    T_round = T_make_dataset + K * (local_epochs * T_epoch + T_predict) + T_update
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def ref_src():
    for cand in (os.path.join(HERE, "_ref", "src"), "/root/reference/src"):
        if os.path.exists(os.path.join(cand, "organization.py")):
            return cand
    return None


def install_shims():
    import torch
    import scipy.sparse._index as _index

    for name in ("matplotlib", "matplotlib.pyplot", "anytree"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    if not getattr(_index.IndexMixin, "_dmt_shim", False):
        orig = _index.IndexMixin.__getitem__

        def conv(k):
            if isinstance(k, torch.Tensor):
                return k.cpu().numpy()
            if isinstance(k, tuple):
                return tuple(conv(x) for x in k)
            return k

        _index.IndexMixin.__getitem__ = lambda self, key: orig(self, conv(key))
        _index.IndexMixin._dmt_shim = True
    if not getattr(torch.load, "_dmt_shim", False):
        orig_load = torch.load

        def load(*a, **kw):
            kw.setdefault("weights_only", False)
            return orig_load(*a, **kw)

        load._dmt_shim = True
        torch.load = load


def enter(src, work, control, device, seed=0, driver="train_recsys_assist"):
    """chdir into a scratch dir with the reference's config.yml + ./data, import the reference driver as a library."""
    import importlib

    install_shims()
    shutil.copy(os.path.join(src, "config.yml"), os.path.join(work, "config.yml"))
    os.chdir(work)
    sys.path.insert(0, src)
    sys.argv = ["ref", "--device", device, "--control_name", control, "--init_seed", str(seed)]
    drv = importlib.import_module(driver)  # argparse runs at import (src/train_recsys_assist.py:21-26)
    import config
    import utils as ref_utils

    ref_utils.process_control()
    cfg = config.cfg
    cfg["seed"] = seed
    cfg["model_tag"] = "{}_{}".format(seed, cfg["control_name"])
    mods = {n: importlib.import_module(n) for n in ("assist", "data", "metrics", "models", "organization", "logger")}
    for n, m in mods.items():
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(src)), (n, m.__file__)
    return types.SimpleNamespace(cfg=cfg, utils=ref_utils, driver=drv, **mods)


def fill_round0(ns, assist, dataset_global):
    """Round-0 containers without running `initialize`: organization_target[0][k] = the global target CSR,
    organization_output[0][k] = per-column mean of the train targets at the same sparsity."""
    import numpy as np
    from scipy.sparse import csr_matrix

    tr = dataset_global["train"].target.tocsr()
    s = np.asarray(tr.sum(0)).ravel()
    c = np.maximum(1, np.diff(tr.tocsc().indptr))
    mean = (s / c).astype(np.float32)
    mean[np.diff(tr.tocsc().indptr) == 0] = float(tr.data.mean())
    for k in ("train", "test"):
        y = dataset_global[k].target.tocsr().astype(np.float32)
        y.sort_indices()
        assist.organization_target[0][k] = y
        assist.organization_output[0][k] = csr_matrix((mean[y.indices], y.indices.copy(), y.indptr.copy()), shape=y.shape)


def run(args):
    src = ref_src()
    if src is None:
        return {"impl": "reference", "unavailable": "baseline/_ref/src is missing (run __graft_entry__.build() where "
                                                    "/root/reference exists)"}
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import dmtcdr_b200  # noqa: F401  (synthetic data writer only; nothing of it is on the timed path)
    from dmtcdr_b200 import synth

    data = synth.make_rating_data(args.data, seed=0)
    work = tempfile.mkdtemp(prefix="dmt_ref_")
    synth.write_reference_layout(data, os.path.join(work, "data"))
    ns = enter(src, work, args.control, args.device)
    cfg = ns.cfg
    threads = args.threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)  # process_control pins 2 (src/utils.py:204); --threads 2 reproduces that
    cfg["local"]["num_epochs"] = 1
    cfg["global"]["num_epochs"] = 1
    torch.manual_seed(0)
    dataset_g = ns.data.fetch_dataset(cfg["data_name"], verbose=False)
    ns.utils.process_dataset(dataset_g)
    data_split = ns.data.split_dataset(dataset_g)
    dataset = ns.data.make_split_dataset(data_split)
    assist = ns.assist.Assist(data_split)
    organization = assist.make_organization()
    names = ["Loss", "RMSE"] if cfg["target_mode"] == "explicit" else ["Loss", "NDCG"]
    metric = ns.metrics.Metric({"train": names, "test": names})
    logger = ns.logger.make_logger(os.path.join(work, "runs"))
    fill_round0(ns, assist, dataset_g)
    K = len(organization)
    nnz_tr = int(dataset_g["train"].target.nnz)
    nnz_te = int(dataset_g["test"].target.nnz)
    sync = (lambda: torch.cuda.synchronize()) if args.device.startswith("cuda") else (lambda: None)

    t0 = time.perf_counter()
    dataset = assist.make_dataset(dataset, 1)
    t_make = time.perf_counter() - t0
    # timed steps: one organization x one local epoch each (round-robin over organizations)
    t_budget0 = time.perf_counter()
    times = []
    done = 0
    why = None
    devnull = open(os.devnull, "w")
    for s in range(args.warmup + args.steps):
        i = s % K
        out_saved = sys.stdout
        sys.stdout = devnull  # the reference prints \r progress lines
        try:
            sync()
            t0 = time.perf_counter()
            organization[i].train(dataset[i]["train"], metric, logger, 1)
            sync()
            dt = time.perf_counter() - t0
        finally:
            sys.stdout = out_saved
        if s >= args.warmup:
            times.append(dt)
            done += 1
        if time.perf_counter() - t_budget0 > args.budget_s and done >= 1 and s + 1 < args.warmup + args.steps:
            why = "time budget {:.0f} s reached after {} timed steps".format(args.budget_s, done)
            break
    logger.reset()
    t_epoch = sum(times) / len(times)
    # predict (train + test) of one organization, update over all organizations' outputs
    out_saved = sys.stdout
    sys.stdout = devnull
    try:
        sync()
        t0 = time.perf_counter()
        out0 = {k: organization[0].predict(dataset[0][k], 1) for k in dataset[0]}
        sync()
        t_pred = time.perf_counter() - t0
        outs = [out0 for _ in range(K)]
        t0 = time.perf_counter()
        assist.update(outs, 1)
        sync()
        t_upd = time.perf_counter() - t0
    finally:
        sys.stdout = out_saved
    L = args.local_epochs
    t_round = t_make + K * (L * t_epoch + t_pred) + t_upd
    visits = K * (L * nnz_tr + nnz_tr + nnz_te)
    shutil.rmtree(work, ignore_errors=True)
    return {"kind": "reference", "device": args.device, "cores": threads, "control_name": args.control,
            "steps_done": done, "warmup": args.warmup, "steps_short_why": why,
            "t_make_dataset_s": t_make, "t_train_epoch_s": t_epoch, "t_predict_s": t_pred, "t_update_s": t_upd,
            "round_s": t_round, "value": visits / t_round, "unit": "rating-visits/s", "visits_per_round": visits,
            "K": K, "local_epochs": L, "nnz_train": nnz_tr, "nnz_test": nnz_te,
            "sample": "unmodified reference (baseline/_ref/src): {} x Organization.train of one organization for one "
                      "local epoch ({:.2f} s each) + make_dataset {:.2f} s + predict of one organization {:.2f} s + "
                      "update {:.2f} s; full round composed as make + K*(L*epoch + predict) + update".format(
                          done, t_epoch, t_make, t_pred, t_upd)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--control", default="ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant")
    ap.add_argument("--data", default="ML1M")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--local-epochs", dest="local_epochs", type=int, default=20)
    ap.add_argument("--budget-s", dest="budget_s", type=float, default=200.0)
    args = ap.parse_args()
    res = run(args)
    print("REF_ARM_JSON " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
