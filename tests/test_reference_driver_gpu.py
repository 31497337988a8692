"""Boundary proof: the UNMODIFIED reference driver src/train_recsys_assist.py, launched through
dmtcdr_b200.launch_reference with the drop-in `models` / `assist` / `organization` ahead of the reference's own, runs a
whole experiment on the GPU — `initialize`, ten assistance rounds (make_dataset -> Organization.train / predict ->
Assist.update), the per-round `test()` loop and the per-round checkpoint of whole objects — and its logged test
metric matches the CPU oracle's replay of the same experiment.

The reference sources are the vendored copy `baseline/_ref/src` (made by __graft_entry__.build(); git-ignored, shipped
to the GPU box) or /root/reference/src in the build container; the test copies them to a scratch directory so the
driver's ./output and ./data stay out of the tree. Nothing of the reference is edited."""
import os
import re
import shutil
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_src():
    for cand in (os.path.join(ROOT, "baseline", "_ref", "src"), "/root/reference/src"):
        if os.path.exists(os.path.join(cand, "train_recsys_assist.py")):
            return cand
    return None


def test_unmodified_driver_runs_on_the_dropin(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    src = _ref_src()
    if src is None:
        pytest.skip("no copy of the reference sources (run __graft_entry__.build() where /root/reference exists)")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import synth

    work = str(tmp_path / "src")
    shutil.copytree(src, work, ignore=shutil.ignore_patterns("__pycache__", "output", "data"))
    data = synth.make_rating_data("tiny-Douban", seed=0)
    synth.write_reference_layout(data, os.path.join(work, "data"))
    control = "Douban_user_explicit_ae_0_genre_assist_constant-0.3_constant"
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), DMT_RNG="reference")
    cmd = [sys.executable, "-m", "dmtcdr_b200.launch_reference", os.path.join(work, "train_recsys_assist.py"),
           "--control_name", control, "--device", "cuda", "--init_seed", "0", "--dmt-compat-shims"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=work)
    log = (out.stdout + out.stderr).replace("\r", "\n")
    assert out.returncode == 0, log[-3000:]
    # the driver resolved the hot-path modules to the drop-in, not to its own src/
    assert "dmtcdr_b200 drop-in active" in log, log[-2000:]
    ckpt = os.path.join(work, "output", "model", "0_{}_checkpoint.pt".format(control))
    assert os.path.exists(ckpt), os.listdir(os.path.join(work, "output", "model"))
    rmse = [float(x) for x in re.findall(r"Test Epoch: \d+\(100%\).*?RMSE: ([0-9.]+)", log)]
    assert len(rmse) >= 10, log[-3000:]
    # the same experiment replayed by the CPU oracle (RNG-identical: the drop-in runs in dmt_rng='reference' mode)
    from oracle import replay

    ref = replay.run_experiment(data, control, seed=0)
    want = ref["metrics"][max(ref["metrics"])]["test/RMSE"]
    assert abs(rmse[-1] - want) <= 2e-4, (rmse[-1], want)  # the driver prints four decimals
