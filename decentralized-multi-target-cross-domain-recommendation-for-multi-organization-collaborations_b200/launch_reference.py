"""Run an UNMODIFIED reference driver with the B200 drop-in modules ahead of the reference's own.

    python -m dmtcdr_b200.launch_reference /path/to/reference/src/train_recsys_assist.py \
        --control_name ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant --device cuda [--dmt-compat-shims]

Python always puts the script's directory first on sys.path, so PYTHONPATH alone cannot shadow the reference's
``models`` / ``assist`` / ``organization``; this launcher builds the path explicitly:
[drop-in dir, reference src, ...], chdir()s into the reference's src (it reads ./config.yml and ./data) and runs the
driver as ``__main__``. ``--dmt-compat-shims`` additionally installs the three environment shims the 2021 reference
needs on a 2025 stack (stub matplotlib/anytree if missing, scipy >= 1.8 fancy-indexing with torch tensors,
``torch.load(weights_only=False)``) — none of them touches the hot path.
"""
import os
import runpy
import sys
import types


def compat_shims():
    import torch
    import scipy.sparse._index as _index

    for name in ("matplotlib", "matplotlib.pyplot", "anytree"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    orig = _index.IndexMixin.__getitem__

    def conv(k):
        if isinstance(k, torch.Tensor):
            return k.cpu().numpy()
        if isinstance(k, tuple):
            return tuple(conv(x) for x in k)
        return k

    _index.IndexMixin.__getitem__ = lambda self, key: orig(self, conv(key))
    orig_load = torch.load
    torch.load = lambda *a, **kw: orig_load(*a, **{"weights_only": False, **kw})


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    driver = os.path.abspath(sys.argv[1])
    args = [a for a in sys.argv[2:] if a != "--dmt-compat-shims"]
    if "--dmt-compat-shims" in sys.argv:
        compat_shims()
    import dmtcdr_b200

    src = os.path.dirname(driver)
    sys.path[:0] = [dmtcdr_b200.DROPIN_DIR, src]
    os.chdir(src)
    sys.argv = [driver] + args
    import assist as _a, models as _m, organization as _o  # noqa: E401  (what the driver's own imports will resolve to)

    here = os.path.abspath(dmtcdr_b200.DROPIN_DIR)
    if not all(os.path.abspath(x.__file__).startswith(here) for x in (_a, _m, _o)):
        raise SystemExit("launch_reference: models / assist / organization did not resolve to the drop-in")
    print("dmtcdr_b200 drop-in active: models, assist, organization <- {}".format(here), flush=True)
    runpy.run_path(driver, run_name="__main__")


if __name__ == "__main__":
    main()
