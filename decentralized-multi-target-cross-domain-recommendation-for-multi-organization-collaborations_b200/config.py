"""Run configuration in the reference's ``cfg`` vocabulary.

The reference keeps one global dict ``config.cfg`` filled from a ``control_name`` string
(reference src/config.py:9-17) and ``utils.process_control()`` (src/utils.py:123-205). When the drop-in
modules run under the reference's own drivers that dict is the source of truth and is used as-is; when
they run standalone (tests on the GPU box, bench.py) :func:`make_cfg` builds an equivalent dict from the
same control string.
"""
from __future__ import annotations

import sys
from collections.abc import MutableMapping

CONTROL_KEYS = ["data_name", "data_mode", "target_mode", "model_name", "info", "data_split_mode", "run_mode", "ar",
                "aw", "match_rate", "pl", "cs"]

_own_cfg: dict = {}


def active_cfg() -> dict:
    """The reference's ``config.cfg`` when its ``config`` module is loaded, else this package's own dict."""
    mod = sys.modules.get("config")
    ref = getattr(mod, "cfg", None) if mod is not None else None
    if isinstance(ref, dict) and "control" in ref:
        return ref
    return _own_cfg


class _CfgProxy(MutableMapping):
    def __getitem__(self, k):
        return active_cfg()[k]

    def __setitem__(self, k, v):
        active_cfg()[k] = v

    def __delitem__(self, k):
        del active_cfg()[k]

    def __iter__(self):
        return iter(active_cfg())

    def __len__(self):
        return len(active_cfg())

    def __contains__(self, k):
        return k in active_cfg()


cfg = _CfgProxy()

_GENRE_ORGS = {"ML100K": 18, "ML1M": 18, "ML10M": 18, "ML20M": 18, "Douban": 3, "Amazon": 4}
_BATCH = {"user": {"ML100K": 100, "ML1M": 500, "ML10M": 1000, "ML20M": 1000, "Douban": 100, "Amazon": 500},
          "item": {"ML100K": 100, "ML1M": 500, "ML10M": 1000, "ML20M": 1000, "Douban": 1000, "Amazon": 500}}


def make_cfg(control_name: str, device: str = "cuda", seed: int = 0, install: bool = True) -> dict:
    """Standalone equivalent of process_args + process_control for one control string
    (field order = reference src/config.yml ``control`` keys; the dict is truncated to the fields given, so
    ``'cs' in cfg`` only when all 12 are present, src/config.py:12-15)."""
    fields = control_name.split("_")
    control = {CONTROL_KEYS[i]: fields[i] for i in range(len(fields))}
    c: dict = {"control": control, "control_name": control_name, "device": device, "num_workers": 0,
               "init_seed": seed, "seed": seed, "num_experiments": 1, "log_interval": 0.25, "world_size": 1,
               "resume_mode": 0, "verbose": False}
    c["data_name"], c["data_mode"], c["target_mode"] = control["data_name"], control["data_mode"], control["target_mode"]
    c["model_name"] = control["model_name"]
    c["info"] = float(control.get("info", 0))
    if "data_split_mode" in control:
        c["data_split_mode"] = control["data_split_mode"]
        if "genre" in c["data_split_mode"]:
            c["num_organizations"] = _GENRE_ORGS[c["data_name"]]
        elif "random" in c["data_split_mode"]:
            c["num_organizations"] = int(c["data_split_mode"].split("-")[1])
        else:
            raise ValueError("Not valid data split mode")
    if "run_mode" in control:
        c["run_mode"] = control["run_mode"]
    c["assist"] = {}
    if "ar" in control and c.get("run_mode") == "assist":
        mode, val = control["ar"].split("-")
        c["assist"]["ar_mode"], c["assist"]["ar"] = mode, float(val)
    if "aw" in control and c.get("run_mode") == "assist":
        c["assist"]["aw_mode"] = control["aw"]
    if "match_rate" in control:
        c["assist"]["match_rate"] = float(control["match_rate"])
    if "pl" in control:
        c["pl"] = control["pl"]
        if c["pl"] != "none":
            mode, val = c["pl"].split("-")
            c["pl_mode"], c["pl_param"] = mode, float(val)
    if "cs" in control:
        c["cs"] = float(control["cs"])
    c["base"] = {}
    c["mf"] = {"hidden_size": 128}
    c["mlp"] = {"hidden_size": [128, 64, 32]}
    c["nmf"] = {"hidden_size": [128, 64, 32]}
    c["ae"] = {"encoder_hidden_size": [256, 128], "decoder_hidden_size": [128, 256]}
    bs = _BATCH[c["data_mode"]][c["data_name"]]
    opt = {"shuffle": {"train": True, "test": False}, "optimizer_name": "Adam", "lr": 1e-3, "betas": (0.9, 0.999),
           "weight_decay": 5e-4, "scheduler_name": "None", "batch_size": {"train": bs, "test": bs}}
    c[c["model_name"]].update(opt)
    c[c["model_name"]]["num_epochs"] = 200 if c["model_name"] != "base" else 1
    c["local"] = dict(opt)
    c["local"]["num_epochs"] = 20
    c["global"] = {"num_epochs": 10}
    c["assist"].update({"optimizer_name": "LBFGS", "lr": 1e-1, "betas": (0.9, 0.999), "weight_decay": 5e-4,
                        "num_epochs": 10})
    c["model_tag"] = "{}_{}".format(seed, control_name)
    c["info_size"] = None
    if install:
        _own_cfg.clear()
        _own_cfg.update(c)
        return _own_cfg
    return c
