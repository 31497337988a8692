"""Harness that imports the UNMODIFIED reference from /root/reference/src in THIS container.

Test infrastructure only (used by make_golden.py to produce the committed fixtures). It never
travels to the GPU box: nothing under tests/ -m gpu, smoke() or bench.py imports this module.

Harness-side shims; no reference file is edited or copied into the repo:
  * stub modules for matplotlib / anytree (imported, unused on the path)
  * scipy index shim so ``csr[:, torch_tensor]`` works on scipy >= 1.8
  * ``torch.load(weights_only=False)`` default
  * cwd = scratch dir holding config.yml (copied at run time) and ./data/<NAME>/processed/*
"""
import os
import shutil
import sys
import types

REF_SRC = "/root/reference/src"


def install_shims():
    import numpy as np
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

        def getitem(self, key):
            return orig(self, conv(key))

        _index.IndexMixin.__getitem__ = getitem
        _index.IndexMixin._dmt_shim = True
    if not getattr(torch.load, "_dmt_shim", False):
        orig_load = torch.load

        def load(*a, **kw):
            kw.setdefault("weights_only", False)
            return orig_load(*a, **kw)

        load._dmt_shim = True
        torch.load = load


def enter(workdir, control_name, seed=0, extra_argv=(), driver_name="train_recsys_assist", first_on_path=()):
    """chdir into ``workdir``, import the reference's modules, parse the control name.
    Returns a namespace with the reference modules and its global cfg."""
    install_shims()
    os.makedirs(workdir, exist_ok=True)
    shutil.copy(os.path.join(REF_SRC, "config.yml"), os.path.join(workdir, "config.yml"))
    os.chdir(workdir)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    for pth in reversed(list(first_on_path)):  # e.g. the drop-in directory, ahead of the reference's own modules
        sys.path.insert(0, pth)
    sys.argv = ["ref", "--device", "cpu", "--control_name", control_name, "--init_seed", str(seed), *extra_argv]
    # the driver builds its argparse from cfg at import time and calls process_args itself
    # (reference src/train_recsys_assist.py:21-26); import it BEFORE anything adds cfg['control_name'].
    import importlib

    driver = importlib.import_module(driver_name)
    import config

    cfg = config.cfg
    import utils as ref_utils

    ref_utils.process_control()
    cfg["seed"] = seed
    cfg["model_tag"] = "{}_{}".format(seed, cfg["control_name"])
    import assist as ref_assist
    import data as ref_data
    import metrics as ref_metrics
    import models as ref_models
    import organization as ref_org
    import logger as ref_logger

    ns = types.SimpleNamespace(cfg=cfg, utils=ref_utils, data=ref_data, models=ref_models, assist=ref_assist,
                               organization=ref_org, metrics=ref_metrics, logger=ref_logger, driver=driver)
    return ns
