#!/usr/bin/env python
"""bench.py — MTAL assist-round throughput (rating-visits/s) on synthetic ML1M-shape data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full assistance round of `ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant`
(BASELINE.json: "MTAL assist-round ratings/sec, ML1M-shape MF/AE"): pseudo-residuals, 18 organizations x 20 local
Adam epochs of the AAE over all 900 188 train residuals, prediction of train+test for every organization, the
exchange of the 18 prediction vectors and the weighted combination. Unit = rating-visit:
K*(20*nnz_train + nnz_train + nnz_test) = 342 071 442 per round.

  value : device-resident rounds (inputs in HBM), CUDA-event timed, max over ranks
  e2e   : the same rounds through the reference-facing drop-in API (Assist.make_dataset / Organization.train /
          predict / Assist.update) with HOST scipy CSR inputs and outputs — uploads and downloads inside the timing
  roofline      : dominant kernel (fused decoder+loss+dZ3) — algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline  : the UNMODIFIED reference (baseline/_ref/src, vendored by __graft_entry__.build()) on a bounded sample,
          all host threads (kind "reference"); the oracle port only if that copy is absent
  --impl reference : the CPU arm alone, same metric/config, --steps/--warmup honoured up to a time budget
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTROL = "ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant"
WORKLOAD = ("ML1M-shape AAE assist round: 6040x3706, 900188 train / 100021 test ratings, 18 genre organizations, "
            "20 local epochs, batch 500 rows")
METRIC = "MTAL assist-round rating-visits/sec (ML1M-shape AAE, K*(20*nnz_train+nnz_train+nnz_test) per round)"
UNIT = "rating-visits/s"
# BASELINE.json configs 3 and 4 (AAE assistance rounds at Douban / Amazon shape; shapes assumed as in SURVEY.md 8d)
OTHER_CONFIGS = [
    ("douban", "Douban_user_explicit_ae_0_genre_assist_constant-0.3_constant", "Douban"),
    ("amazon", "Amazon_user_implicit_ae_0_genre_assist_constant-0.1_optim_0.5_dp-10", "Amazon"),
]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense bf16 TFLOP/s (burst figure: the kernel is timed alone)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["bf16_tflops"]), "measured bf16 burst (MEASURED_PEAKS.json)"
    return 1590.0, "fallback (B200_PROFILING.md)"


NCU_RAW = os.path.join(ROOT, "profiles", "r2_ncu_raw_fused.csv")
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def ncu_per_kernel(path=NCU_RAW):
    """Per-kernel averages parsed from the committed `ncu --set full` raw page (scripts/r2_prof.sh writes it):
    {kernel: {'dram_bytes', 'l2_to_sm_bytes', 'us', 'regs', 'launches'}}. Raises when the file is missing: the
    roofline block's `traffic` is measured evidence or nothing."""
    import csv
    import re

    if not os.path.exists(path):
        raise FileNotFoundError("{} is missing: run scripts/r2_prof.sh under gpurun and commit its raw page".format(path))
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        i = col[key]
        return float(r[i].replace(",", "")) * _UNIT.get(units[i], 1.0)

    acc = {}
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        ids = re.findall(r"([A-Za-z_]\w*)(?=[<(])", r[col["Kernel Name"]])
        name = next((x for x in ids if x.endswith("_kernel")), ids[0] if ids else r[col["Kernel Name"]])
        a = acc.setdefault(name, {"dram_bytes": 0.0, "l2_to_sm_bytes": 0.0, "us": 0.0, "regs": 0, "launches": 0})
        a["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        a["l2_to_sm_bytes"] += val(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
        a["us"] += val(r, "gpu__time_duration.sum")
        a["regs"] = int(float(r[col["launch__registers_per_thread"]]))
        a["launches"] += 1
    for a in acc.values():
        for k in ("dram_bytes", "l2_to_sm_bytes", "us"):
            a[k] /= a["launches"]
    return acc


# profile_step class -> kernel of csrc/fused.cu (bulk gather mode / register-load mode)
CLASS_KERNEL = {"fwd_rows": ("ae_fwd_rows_kernel",) * 2, "decoder_loss_dz3": ("ae_dec_chunks_bulk_kernel", "ae_dec_chunks4_kernel"),
                "dw4_segments": ("ae_seg_chunks_bulk_kernel", "ae_seg_chunks_kernel"), "bwd_rows": ("ae_bwd_rows_kernel",) * 2,
                "grad_phase": ("ae_grad_phase_kernel",) * 2, "grad_norm": ("norm_prepare_kernel",) * 2,
                "clip_adam": ("adam_shadow_kernel",) * 2}


def step_algorithmic_bytes(t_batch, d_batch, n_seg_t, n_seg_d, B, n_params, H1=256, H2=128):
    """Compulsory bytes per launch of every kernel of the fused step (SURVEY.md section 8d's per-unit figures x the
    units of one batch): gathered 1 KB rows count once per entry, dense operands once."""
    act = 4 * B
    wts = 4 * 2 * H1 * H2
    return {
        "fwd_rows": d_batch * (4 * H1 + 8) + wts + act * (H1 + H2 + H2 + H1),
        "decoder_loss_dz3": t_batch * (4 * H1 + 16) + act * H1 * 2,
        "dw4_segments": t_batch * (4 * H1 + 8) + n_seg_t * (4 * H1 + 4),
        "bwd_rows": wts + act * (H1 + H2 + H1 + H2 + H1),
        "grad_phase": d_batch * (4 * H1 + 8) + n_seg_d * 4 * H1 + act * (H1 + H2 + H2 + H1) + wts,
        "grad_norm": 4 * n_params,
        "clip_adam": 28 * n_params + 4 * n_params,
    }


def step_roofline(prof, bytes_by_class, gather_mode, hbm, peak_src):
    """The roofline block: every kernel class of the step with its algorithmic bytes, CUDA-event time and the DRAM /
    L2->SM traffic ncu measured for that kernel (parsed from profiles/, per launch); the headline entry is the class
    with the longest launch."""
    ncu = ncu_per_kernel()
    idx = 0 if gather_mode == "bulk" else 1
    classes = {}
    for cls, ms in prof.items():
        kern = CLASS_KERNEL[cls][idx]
        n = ncu.get(kern)
        b = bytes_by_class[cls]
        ach = b / (ms * 1e-3) / 1e9
        classes[cls] = {"kernel": kern, "ms_per_launch": ms, "algorithmic_bytes_per_launch": b, "achieved_GBps": ach,
                        "frac_of_hbm_peak": ach / hbm,
                        "traffic_dram_bytes": n["dram_bytes"] if n else None,
                        "traffic_l2_to_sm_bytes": n["l2_to_sm_bytes"] if n else None,
                        "ncu_us": n["us"] if n else None, "regs": n["regs"] if n else None}
    dom = max(prof, key=prof.get)
    d = classes[dom]
    return {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved_GBps"], "peak": hbm, "unit": "GB/s",
            "frac": d["frac_of_hbm_peak"], "traffic": d["traffic_dram_bytes"],
            "traffic_source": "profiles/r2_ncu_raw_fused.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum "
                              "of this kernel, per launch, parsed at bench time)",
            "peak_source": peak_src, "ms_per_launch": d["ms_per_launch"],
            "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
            "note": "ML1M-shape tables (W4 3.8 MB per organization) are L2-resident: the algorithmic bytes are served by "
                    "L2 (traffic_l2_to_sm_bytes), DRAM traffic is far lower; hbm_case runs the same gather arithmetic "
                    "on a shape L2 cannot hold",
            "slowest_class": dom, "step_classes": classes, "step_sum_us": 1e3 * sum(prof.values())}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe's clocks line)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "samples": len(self.rows), "reasons": reasons}


def build_problem(seed=0):
    """Synthetic ML1M-shape data + the genre split, identical on every rank."""
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import runner, synth
    from dmtcdr_b200.config import cfg, make_cfg

    make_cfg(CONTROL, device="cuda" if torch.cuda.is_available() else "cpu", seed=seed)
    data = synth.make_rating_data("ML1M", seed=0)
    torch.manual_seed(seed)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    data_split = runner.split_dataset(dataset)
    mats = {k: (dataset[k].data, dataset[k].target) for k in dataset}
    return data, dataset, data_split, mats, cfg


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_round_sample(mats, data_split, n_orgs_sample=1, epochs_sample=1, batch_rows=500, threads=None):
    """Oracle port (CPU restatement of the reference) timed on a bounded sample of the SAME workload:
    `n_orgs_sample` organizations x `epochs_sample` local epochs (+ their predict), plus make_dataset and update for
    all organizations; returns (rating-visits processed, seconds)."""
    from oracle import mtal, replay, train

    if threads:
        torch.set_num_threads(threads)
    y = {k: mats[k][1] for k in mats}
    cols = [s.numpy() for s in data_split]
    K = len(cols)
    n_rows, n_cols = y["train"].shape
    F0 = {k: np.full(y[k].nnz, 3.5, np.float32) for k in y}
    t0 = time.perf_counter()
    res = {k: mtal.residual(F0[k], y[k].data, "explicit", False) for k in y}
    tgt = {k: replay._with_data(y[k], res[k]) for k in y}
    visits = 0
    outs = []
    for i in range(n_orgs_sample):
        data_i = mats["train"][0][:, cols[i]].tocsr()
        p0 = replay.init_ae_params(data_i.shape[1], n_cols)
        epoch_batches, masks = [], []
        for _ in range(epochs_sample):
            batches = replay.loader_batches(n_rows, batch_rows, True)
            epoch_batches.append(batches)
            masks += [replay.draw_keep_mask(len(b)) for b in batches]
        p, _ = train.train_org_ae(p0, data_i, tgt["train"], "user", "explicit", epoch_batches, masks)
        visits += epochs_sample * y["train"].nnz
        o = {k: train.predict_org_ae(p, data_i, tgt[k], "user", "explicit", batch_rows) for k in y}
        visits += y["train"].nnz + y["test"].nnz
        outs.append(o)
    org_out = [outs[i % len(outs)] for i in range(K)]
    mtal.update(F0, {k: y[k].data for k in y}, org_out, {k: y[k].indices for k in y}, cols, n_cols, "explicit", 0.1)
    return visits, time.perf_counter() - t0


def reference_leg(threads, steps, warmup, device="cpu", budget_s=200.0, timeout=900, control=None, data="ML1M"):
    """Time the UNMODIFIED reference (baseline/_ref/src) in a child process (its global cfg / argparse-at-import do not
    mix with this process): baseline/ref_arm.py. Returns its result dict, or {'unavailable': why}."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_arm.py"), "--control", control or CONTROL, "--data", data,
           "--device", device, "--threads", str(threads), "--steps", str(steps), "--warmup", str(warmup),
           "--budget-s", str(budget_s)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"unavailable": "reference leg timed out after {} s".format(timeout)}
    for line in out.stdout.splitlines():
        if line.startswith("REF_ARM_JSON "):
            return json.loads(line[len("REF_ARM_JSON "):])
    return {"unavailable": "reference leg failed: " + (out.stderr.strip().splitlines() or ["no output"])[-1][:300]}


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path (unmodified sources vendored by build()
    to baseline/_ref, driven through Assist.make_dataset / Organization.train / predict / Assist.update), all host
    threads, same metric / config. A step = Organization.train of one organization for one local epoch; the round
    figure composes the timed pieces as the reference's loop does (baseline/ref_arm.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    res = reference_leg(threads, args.steps, args.warmup, budget_s=240.0)
    if "unavailable" in res:
        # the vendored copy is absent (build() ran where /root/reference does not exist): fall back to the oracle port
        _, _, data_split, mats, _ = build_problem()
        v, dt = cpu_round_sample(mats, data_split, 1, 20, threads=threads)
        res_port = {"value": v / dt, "round_s": None, "kind": "port", "cores": threads, "steps_done": 1,
                    "sample": "oracle port, 1 of 18 organizations x 20 local epochs + predict + residual/update "
                              "(reference copy unavailable: {})".format(res["unavailable"])}
        res = res_port
    two = reference_leg(2, 2, 1, budget_s=60.0) if res.get("kind") == "reference" else None
    value = res["value"]
    ms = 1e3 * res["round_s"] if res.get("round_s") else 1e3 * (18 * (20 * 900188 + 1000209)) / value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": res.get("steps_done", args.steps), "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "control_name": CONTROL},
            "steps_requested": args.steps, "steps_short_why": res.get("steps_short_why"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "reference_2_threads": None if not two or "unavailable" in two else {
                "value": two["value"], "unit": UNIT, "cores": 2, "round_s": two["round_s"],
                "note": "torch.set_num_threads(2) as the reference pins it (src/utils.py:204)", "sample": two["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    from dmtcdr_b200 import dist as D

    # Libraries (NCCL's version banner, ...) print to the C-level stdout; the contract is ONE JSON line there. Everything
    # before the final print is redirected to stderr at the file-descriptor level.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:{}".format(local)
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import engine as E
    from dmtcdr_b200 import native, roundloop, runner
    from dmtcdr_b200.config import cfg

    native.load()
    data, dataset, data_split, mats, _ = build_problem()
    rounds = roundloop.AssistRounds(mats, [s.numpy() for s in data_split], "explicit", 500, clamp=False, ar=0.1,
                                    local_epochs=args.local_epochs, rank=rank, world=world, device=dev)
    rounds.round0()
    exchange = (lambda O: D.exchange_outputs(rounds.state.O_full, rounds.chunk, rank, world)) if world > 1 else None
    visits = rounds.rating_visits_per_round()

    def one_round(t):
        rounds.run_round(t, exchange)

    # clocks / throttle reasons are sampled from the last warm-up round on (same load as the timed rounds), so that even a
    # short timed region (8 GPUs: ~40 ms per round) has samples taken under load
    sampler = ClockSampler(local) if rank == 0 else None
    for w in range(args.warmup):
        if sampler and w == args.warmup - 1:
            sampler.start()
        one_round(w + 1)
    rounds.sync()
    D.barrier()
    torch.cuda.synchronize()
    if sampler and not sampler.is_alive():
        sampler.start()
    launches0 = native.load().dmt_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        one_round(args.warmup + s + 1)
    # the current stream already waits for every organization stream (signal_current in run_round) and the update
    ev1.record()
    rounds.sync()
    D.barrier()
    torch.cuda.synchronize()
    ms_total = D.max_over_ranks(ev0.elapsed_time(ev1), dev)
    launches = native.load().dmt_launch_count() - launches0
    clocks = sampler.summary() if sampler else None
    ms_step = ms_total / args.steps
    value = visits / (ms_step / 1e3)

    # ---- roofline: every kernel class of the step, CUDA events on the launching stream (dmt_org_profile_step), on the
    #      grid the timed rounds use; ncu traffic per kernel parsed from profiles/ (fails loudly when absent)
    roof = None
    if rank == 0:
        org = rounds.my_orgs[0]
        eng = rounds.eng[org]
        prof = eng.h.profile_step(b=0, reps=20)
        n_params = eng.h.n_params
        batches = E.fast_perm_batches(rounds.n_rows, 500)[:-1]
        tl, dl = eng.t_len, eng.d_len
        t_batch = float(np.mean([tl[b].sum() for b in batches]))
        d_batch = float(np.mean([dl[b].sum() for b in batches]))
        y_host = mats["train"][1]
        d_host = mats["train"][0][:, data_split[org].numpy()].tocsr()
        rows0 = np.sort(batches[0])
        n_seg_t = int(np.unique(y_host[rows0].indices).size)
        n_seg_d = int(np.unique(d_host[rows0].indices).size)
        hbm, peak_src = measured_peaks()
        by_class = step_algorithmic_bytes(t_batch, d_batch, n_seg_t, n_seg_d, 500, n_params)
        roof = step_roofline(prof, by_class, eng.h.gather_mode(), hbm, peak_src)
        roof["grid_note"] = "profiled on the launch configuration of the timed rounds ({} organizations on this GPU)".format(
            len(rounds.my_orgs))
        # whole-round aggregate: compulsory bytes of one round / the CUDA-event time of the timed rounds
        agg_bytes = E.bytes_per_round(rounds.K, rounds.state.y["train"].nnz, rounds.state.y["test"].nnz, rounds.n_rows,
                                      [len(c) for c in rounds.cols], rounds.state.n_cols, args.local_epochs, 500)
        roof["round_aggregate"] = {"algorithmic_bytes_per_round": agg_bytes, "ms_per_round": ms_step,
                                   "achieved_GBps": agg_bytes / (ms_step * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": agg_bytes / (ms_step * 1e-3) / 1e9 / (hbm * world),
                                   "note": "sum over all organizations and steps of the per-kernel compulsory bytes "
                                           "(engine.bytes_per_round) over the measured round time, against N x the "
                                           "measured HBM copy peak; most of it is served by L2 at this shape"}
        roof["hbm_case"] = hbm_bound_case(dev, hbm)
        # the decoder's other form (tcgen05 GEMMs, 3xTF32): same batch, same plan, per-class timings + D1 alone
        eng.set_decoder("tc")
        prof_tc = eng.h.profile_step(b=0, reps=20)
        eng.set_decoder(E.decoder_default() if eng.target.sorted else "gather")
        roof["tc_decoder"] = tc_decoder_case(rounds.state.y["train"], dev, prof, prof_tc)

    # ---- end to end through the drop-in API with host buffers (single-process API: measured on rank 0's GPU)
    e2e = run_e2e(args, data, rank, world, dev)
    if world == 1:  # the same rounds with every organization's state_dict read back each round, as the reference does
        eager = run_e2e(args, data, rank, world, dev, state_dicts=True, n_steps=2)
        e2e["e2e_with_state_dicts"] = {k: eager[k] for k in ("value", "ms_per_step", "d2h_bytes_per_step", "state_dicts")}

    # the side blocks below run in a clean state: the headline's engines (18 organizations' parameters, plans and
    # graphs) and whatever the e2e experiments left in torch's caching allocator are released first
    line_cfg = {"orgs_per_rank": len(rounds.my_orgs),
                "decoder": rounds.eng[rounds.my_orgs[0]].decoder if rounds.my_orgs else None,
                "step_fanout": bool(rounds.fanout)}
    if world == 1:
        import gc
        rounds.close()
        gc.collect()
        torch.cuda.empty_cache()
    mf = run_mf_joint(data, dev) if (rank == 0 and world == 1) else None
    nmf = None
    if rank == 0 and world == 1 and args.configs != "none":
        try:
            nmf = run_nmf_alone(data, dev)
        except Exception as e:
            nmf = {"error": "{}: {}".format(type(e).__name__, e)}
    # the other BASELINE.json configurations, as extra blocks of the line (N=1 only; each with its own CPU leg)
    more = None
    if rank == 0 and world == 1 and args.configs != "none":
        more = {}
        for key, control, data_name in OTHER_CONFIGS:
            if args.configs in ("all", key):
                try:
                    more[key] = run_config_block(control, data_name, dev)
                except Exception as e:  # a failing side block must not take the headline line with it
                    more[key] = {"control_name": control, "error": "{}: {}".format(type(e).__name__, e)}
    if world > 1:
        D.barrier()
        import torch.distributed as tdist
        tdist.destroy_process_group()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cpu = None
    if world == 1:  # the CPU baseline is reported on rank 0 at N=1 only
        ref = reference_leg(threads, 3, 1, budget_s=60.0)
        if "unavailable" not in ref:
            cpu = {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": "reference",
                   "round_s": ref["round_s"], "sample": ref["sample"]}
            # informative: the same unmodified reference with --device cuda (PyTorch eager on this B200)
            eager = reference_leg(threads, 2, 1, device="cuda", budget_s=60.0)
            cpu["reference_cuda_eager"] = eager if "unavailable" in eager else {
                "value": eager["value"], "unit": UNIT, "round_s": eager["round_s"], "sample": eager["sample"]}
        else:
            v, dt = cpu_round_sample(mats, data_split, 1, 20, threads=threads)
            cpu = {"value": v / dt, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "1 of 18 organizations x all 20 local epochs + its predict + residual/update for all "
                             "organizations ({:.1f} s of CPU work; oracle/ torch-CPU port; reference copy unavailable: "
                             "{})".format(dt, ref["unavailable"])}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "control_name": CONTROL, "local_epochs": args.local_epochs,
                       "orgs_per_rank": line_cfg["orgs_per_rank"], "parallelism": "org-sharded x{}".format(world),
                       "decoder": line_cfg["decoder"], "step_fanout": line_cfg["step_fanout"],
                       "l2_policy": "inputs larger than L2: per-round working set (18 x 17 MB parameters+moments, "
                                    "2 x 72 MB prediction matrices, plans) exceeds 126 MB"},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roof,
            "cpu_baseline": cpu, "mf_joint": mf, "nmf_alone": nmf, "other_configs": more}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


def hbm_bound_case(dev, hbm):
    """The SAME decoder kernel (ae_decoder_fwd_kernel, train mode) on a shape that cannot be served by the 126 MB L2:
    512 batch rows x 2048 targets each, every target a DISTINCT column of a 1 100 000 x 256 fp32 weight matrix
    (1.13 GB), so every 1 KB weight row is read from HBM exactly once. Raw C-ABI call, outputs preallocated, CUDA
    events on the launching stream."""
    from dmtcdr_b200 import native

    lib = native.load()
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    n_rows, n_dec, per_row, H = 512, 1_100_000, 2048, 256
    nnz = n_rows * per_row
    cols = torch.randperm(n_dec, device=dev, generator=g)[:nnz].reshape(n_rows, per_row)
    cols, _ = torch.sort(cols, dim=1)
    indptr = (torch.arange(n_rows + 1, device=dev, dtype=torch.int64) * per_row).to(torch.int32).contiguous()
    indices = cols.reshape(-1).to(torch.int32).contiguous()
    target = torch.randn(nnz, device=dev, generator=g)
    A3 = torch.tanh(torch.randn(n_rows, H, device=dev, generator=g))
    W4 = torch.empty(n_dec, H, device=dev).normal_(0, 0.05, generator=g)
    b4 = torch.zeros(n_dec, device=dev)
    rows = torch.arange(n_rows, device=dev, dtype=torch.int32)
    pred = torch.empty(nnz, device=dev)
    gout = torch.empty(nnz, device=dev)
    dz3 = torch.empty(n_rows, H, device=dev)
    loss_rows = torch.empty(n_rows, device=dev)
    n_t = torch.tensor([nnz], device=dev, dtype=torch.int32)
    P = native.ptr

    def launch():
        native.check(lib.dmt_ae_decoder_fwd(P(rows), n_rows, P(indptr), P(indices), P(target), P(A3), P(W4), P(b4), H, 0,
                                            P(n_t), P(pred), P(gout), P(dz3), P(loss_rows), 1, native.stream()),
                     "dmt_ae_decoder_fwd")

    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # per target: 1 KB weight row + 4 B column + 4 B target + 4 B bias read, 4 B prediction + 4 B gradient written;
    # per row: A3 in, dZ3 out
    bytes_ = nnz * (4 * H + 20) + n_rows * H * 4 * 2
    ach = bytes_ / (ms * 1e-3) / 1e9
    return {"kernel": "ae_decoder_fwd_kernel<2>", "shape": "512 rows x 2048 distinct targets over 1.1M columns (W4 1.13 GB >> L2)",
            "ms_per_launch": ms, "algorithmic_bytes_per_launch": bytes_, "achieved": ach, "peak": hbm,
            "frac": ach / hbm, "unit": "GB/s",
            "traffic": ncu_per_kernel(os.path.join(ROOT, "profiles", "r1_ncu_raw_decoder_hbm.csv"))[
                "ae_decoder_fwd_kernel"]["dram_bytes"],
            "traffic_source": "profiles/r1_ncu_raw_decoder_hbm.csv (ncu --set full of this kernel on this shape: dram read + "
                              "write per launch, parsed at bench time)"}


def tc_decoder_case(y_csr, dev, prof_gather, prof_tc):
    """The tensor-core form of the decoder's last layer (csrc/decoder_tc.cu; tcgen05.mma kind::tf32, 3xTF32 parity
    split) on one real ML1M-shape batch — the first 500 rows of the train target CSR: the forward GEMM + masked
    epilogue (D1) alone through the raw C-ABI with preallocated outputs, CUDA events on the launching stream. Tensor
    roofline: ALGORITHMIC flops 2*B*N*H (one fp32 product; the three TF32 passes are the price of parity, not work)
    against the measured dense bf16 peak; TF32 runs at half the bf16 rate, so 1/6 of that peak is this kernel's
    ceiling. Next to it the engine's per-class step timings in both decoder modes (same batch, same plan)."""
    from dmtcdr_b200 import native

    lib = native.load()
    n_rows, H = 500, 256
    n_dec = y_csr.shape[1]
    nnz = int(y_csr.indptr_host[n_rows])
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    A3 = torch.tanh(torch.randn(n_rows, H, device=dev, generator=g))
    W4 = torch.empty(n_dec, H, device=dev).normal_(0, 0.05, generator=g)
    b4 = torch.zeros(n_dec, device=dev)
    rows = torch.arange(n_rows, device=dev, dtype=torch.int32)
    pred = torch.empty(y_csr.nnz, device=dev)
    scratch = torch.empty(lib.dmt_ae_decoder_tc_scratch_floats(n_rows, n_dec, H), device=dev)
    P = native.ptr

    def launch(passes):
        native.check(lib.dmt_ae_decoder_tc(P(rows), n_rows, P(y_csr.indptr), P(y_csr.indices), None, P(A3), P(W4), P(b4),
                                           H, n_dec, 0, None, passes, P(pred), None, None, None, None, None, 0,
                                           P(scratch), native.stream()), "dmt_ae_decoder_tc")

    out = {}
    for passes in (3, 1):
        for _ in range(3):
            launch(passes)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            launch(passes)
        e1.record()
        torch.cuda.synchronize()
        out[passes] = e0.elapsed_time(e1) / reps
    flops = 2.0 * n_rows * n_dec * H
    peak, src = measured_tensor_peak()
    ach = flops / (out[3] * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "tile_tab_kernel + dec_fwd_tc_kernel (D1: A3 W4^T on tcgen05, masked epilogue)",
            "shape": "{} rows x {} items x H {} ({} targets, {:.1f} % dense)".format(n_rows, n_dec, H, nnz,
                                                                                   100.0 * nnz / (n_rows * n_dec)),
            "ms_per_launch_3xTF32": out[3], "ms_per_launch_1xTF32": out[1],
            "algorithmic_flops_per_launch": flops, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": src,
            "tensor_pipe_source": "profiles/r1_ncu_raw_tc.csv: sm__pipe_tensor_cycles_active 20 % (D1), 13 % (D2), "
                                  "18 % (D3) of peak sustained active",
            "step_kernel_ms_gather": {k: prof_gather[k] for k in ("decoder_loss_dz3", "dw4_segments")},
            "step_kernel_ms_tc": {k: prof_tc[k] for k in ("decoder_loss_dz3", "dw4_segments")},
            "verdict": "measured negative at all three BASELINE shapes (other_configs.*.tc_vs_gather): the gather form "
                       "is the engine default, the tensor-core form stays opt-in (DMT_DECODER=tc)",
            "note": "latency-bound at this size (116 CTAs x 8 k-chunks); the gather form stays the engine default at "
                    "ML1M shape (DESIGN.md section 5)"}


def run_mf_joint(data, dev, epochs=3):
    """Config 1 (`ML1M_user_explicit_mf_0_genre_joint`, reference src/train_recsys_joint.py:118-134): joint MF epochs
    through the drop-in `models.mf` (fused gather+dot+bias+loss kernel, sort + segmented-reduction gradients) with the
    driver-owned clip_grad_norm_ + torch.optim.Adam, batches of 500 users resident on the device. Returns ratings/s,
    the forward kernel's roofline and the oracle port on the host cores for one epoch."""
    import dmtcdr_b200
    from dmtcdr_b200 import native, runner
    from dmtcdr_b200.config import cfg, make_cfg

    make_cfg("ML1M_user_explicit_mf_0_genre_joint", device="cuda", seed=0)
    models, _, _ = dmtcdr_b200.use_dropin()
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    ds = dataset["train"]
    n_rows = len(ds)
    torch.manual_seed(0)
    model = models.mf().to(dev)
    model.train(True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=5e-4)
    g = torch.Generator().manual_seed(0)

    def epoch_batches():
        perm = torch.randperm(n_rows, generator=g).numpy()
        return [{k: v.to(dev) for k, v in runner.pair_batch(ds, perm[s:s + 500]).items()} for s in range(0, n_rows, 500)]

    def run_epoch(batches):
        for b in batches:
            opt.zero_grad()
            out = model(b)
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
            opt.step()
        return out["loss"]

    run_epoch(epoch_batches())  # warm-up
    sets = [epoch_batches() for _ in range(epochs)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for bs in sets:
        loss = run_epoch(bs)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / epochs
    nnz = ds.data.nnz
    # forward kernel alone on one batch: 2 embedding rows (4*128 B) + 2 biases + 2 indices + rating + pred + dpred
    b = sets[0][0]
    u32, i32 = b["user"].to(torch.int32), b["item"].to(torch.int32)
    Wu, Wi = model.user_weight.weight.detach(), model.item_weight.weight.detach()
    bu, bi = model.user_bias.weight.detach().view(-1), model.item_bias.weight.detach().view(-1)
    for _ in range(3):
        native.mf_fwd(u32, i32, b["rating"], Wu, Wi, bu, bi, model.bias.detach(), 0)
    e0.record()
    for _ in range(20):
        native.mf_fwd(u32, i32, b["rating"], Wu, Wi, bu, bi, model.bias.detach(), 0)
    e1.record()
    torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / 20
    n_b = u32.numel()
    bytes_fwd = n_b * (2 * 4 * 128 + 2 * 4 + 2 * 4 + 4 + 4 + 4)
    hbm, _ = measured_peaks()
    # CPU: the oracle port, one epoch
    from oracle import models as om
    from oracle import train as otrain
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    oopt = otrain.Adam(sd)
    perm = torch.randperm(n_rows, generator=g).numpy()
    t0 = time.perf_counter()
    for s in range(0, n_rows, 500):
        cb = runner.pair_batch(ds, perm[s:s + 500])
        otrain.train_step(oopt, lambda p: om.pair_forward("mf", p, cb, "explicit", True))
    cpu_sec = time.perf_counter() - t0
    return {"workload": "ML1M_user_explicit_mf_0_genre_joint: joint MF epoch, 900188 ratings, 13 batches of 500 users, "
                        "1257235 parameters, driver-owned clip + torch.optim.Adam",
            "value": nnz / sec, "unit": "ratings/s", "ms_per_epoch": 1e3 * sec, "last_loss": float(loss),
            "mf_fwd_kernel": {"ms_per_launch": ms_fwd, "ratings": n_b, "algorithmic_bytes_per_launch": bytes_fwd,
                              "achieved_GBps": bytes_fwd / (ms_fwd * 1e-3) / 1e9,
                              "frac_of_hbm_peak": bytes_fwd / (ms_fwd * 1e-3) / 1e9 / hbm,
                              "note": "tables (3 MB + 1.9 MB) are L2-resident at ML1M shape"},
            "cpu_baseline": {"value": nnz / cpu_sec, "unit": "ratings/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "one joint-MF epoch with the oracle port ({:.1f} s)".format(cpu_sec)}}


def run_nmf_alone(data, dev, epochs=2):
    """Config 2 (`ML1M_item_implicit_nmf_0_random-8_alone`; the literal `nmf_1` control crashes in the reference on
    ML1M-shape data, SURVEY.md section 8c, so `info=0` is what is timed): 8 organizations, each training its own NCF
    (drop-in `models.nmf`: embedding gather + concat kernel, FFMA tower, GMF product + affine + BCE fused) on its block of
    user columns with the driver-owned clip_grad_norm_ + torch.optim.Adam (reference src/train_recsys_alone.py:130-147),
    batches of 500 items resident on the device. ratings/s over all organizations + the oracle port on the host cores for
    one organization-epoch."""
    import dmtcdr_b200
    from dmtcdr_b200 import runner
    from dmtcdr_b200.config import make_cfg

    control = "ML1M_item_implicit_nmf_0_random-8_alone"
    cfg = make_cfg(control, device="cuda", seed=0)
    models, _, _ = dmtcdr_b200.use_dropin()
    torch.manual_seed(0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    split = runner.split_dataset(dataset)
    org_ds = runner.make_split_dataset(dataset, split)
    K = len(org_ds)
    bs = cfg["nmf"]["batch_size"]["train"] if "batch_size" in cfg["nmf"] else 500
    org_models, opts = [], []
    for k in range(K):
        ds = org_ds[k]["train"]
        m = models.nmf(ds.num_users["data"], ds.num_items["data"]).to(dev)
        m.train(True)
        org_models.append(m)
        opts.append(torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=5e-4))
    g = torch.Generator().manual_seed(0)

    def epoch_batches(k):
        ds = org_ds[k]["train"]
        perm = torch.randperm(len(ds), generator=g).numpy()
        out = []
        for s0 in range(0, len(ds), bs):
            b = runner.pair_batch(ds, perm[s0:s0 + bs])
            if len(b[cfg["data_mode"]]) > 0:
                out.append({kk: v.to(dev) for kk, v in b.items()})
        return out

    def run_epoch(k, batches):
        loss = None
        for b in batches:
            opts[k].zero_grad()
            out = org_models[k](b)
            out["loss"].backward()
            torch.nn.utils.clip_grad_norm_(org_models[k].parameters(), 1)
            opts[k].step()
            loss = out["loss"]
        return loss

    for k in range(K):
        run_epoch(k, epoch_batches(k))  # warm-up
    sets = [[epoch_batches(k) for k in range(K)] for _ in range(epochs)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for per_org in sets:
        for k in range(K):
            loss = run_epoch(k, per_org[k])
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / epochs
    nnz = sum(org_ds[k]["train"].data.nnz for k in range(K))
    n_steps = sum(len(per_org[k]) for per_org in sets[:1] for k in range(K))
    # CPU: the oracle port, organization 0, one epoch
    from oracle import models as om
    from oracle import train as otrain
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {kk: v.detach().cpu() for kk, v in org_models[0].state_dict().items()}
    oopt = otrain.Adam(sd)
    ds0 = org_ds[0]["train"]
    perm = torch.randperm(len(ds0), generator=g).numpy()
    t0 = time.perf_counter()
    for s0 in range(0, len(ds0), bs):
        cb = runner.pair_batch(ds0, perm[s0:s0 + bs])
        if len(cb[cfg["data_mode"]]) > 0:
            otrain.train_step(oopt, lambda p: om.pair_forward("nmf", p, cb, "implicit", True))
    cpu_sec = time.perf_counter() - t0
    return {"workload": "{}: {} organizations (random user blocks), {} train ratings in total, batches of {} items, "
                        "{} optimizer steps per epoch over all organizations".format(control, K, nnz, bs, n_steps),
            "value": nnz / sec, "unit": "ratings/s", "ms_per_epoch_all_orgs": 1e3 * sec,
            "us_per_step": 1e6 * sec / max(n_steps, 1), "last_loss": float(loss),
            "note": "module-level path: the driver owns backward / clip / Adam, so every step pays torch's eager "
                    "dispatch around the kernels (DESIGN.md: host-bound at this size)",
            "cpu_baseline": {"value": ds0.data.nnz / cpu_sec, "unit": "ratings/s", "cores": os.cpu_count() or 1,
                             "kind": "port", "sample": "organization 0, one epoch with the oracle port ({:.1f} s)".format(
                                 cpu_sec)}}


def run_config_block(control, data_name, dev, n_rounds=2, n_warm=2, local_epochs=20, cpu_leg=True):
    """One more BASELINE.json configuration as a block of the bench line: device-resident assistance rounds
    (roundloop.AssistRounds, same code path as `value`) on the synthetic shape SURVEY.md section 8d fixes for it, CUDA-event
    timed after warm-up; the per-class step profile of one organization; the decoder kernel's algorithmic-byte rate;
    and the unmodified reference on the host cores for the same control string (bounded sample)."""
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import engine as E
    from dmtcdr_b200 import roundloop, runner, synth
    from dmtcdr_b200.config import make_cfg

    cfg = make_cfg(control, device="cuda", seed=0)
    data = synth.make_rating_data(data_name, seed=0)
    torch.manual_seed(0)
    dataset = runner.fetch_dataset(data)
    runner.process_dataset(dataset)
    split = [s.numpy() for s in runner.split_dataset(dataset)]
    mats = {k: (dataset[k].data, dataset[k].target) for k in dataset}
    a = cfg["assist"]
    clamp = cfg["data_name"] in ("Douban", "Amazon") and not (
        cfg["data_name"] == "Douban" and cfg["data_mode"] == "item" and cfg["target_mode"] == "explicit")
    privacy = (cfg["pl_mode"], cfg["pl_param"]) if cfg.get("pl", "none") != "none" else None
    bs = cfg["local"]["batch_size"]["train"]
    R = roundloop.AssistRounds(mats, split, cfg["target_mode"], bs, clamp=clamp, ar=a["ar"], ar_mode=a["ar_mode"],
                               aw_mode=a["aw_mode"], match_rate=a.get("match_rate", 1.0), local_epochs=local_epochs,
                               device=dev, privacy=privacy)
    R.round0()
    for t in range(1, n_warm + 1):
        R.run_round(t)
    R.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"train_predict": 0.0, "update": 0.0}
    e0.record()
    for t in range(n_warm + 1, n_warm + n_rounds + 1):
        h0 = time.perf_counter()
        R.train_predict(t)
        h1 = time.perf_counter()
        R.combine()
        phase["train_predict"] += h1 - h0
        phase["update"] += time.perf_counter() - h1
    e1.record()
    R.sync()
    ms = e0.elapsed_time(e1) / n_rounds
    visits = R.rating_visits_per_round()
    eng = R.eng[R.my_orgs[0]]
    prof = eng.h.profile_step(b=0, reps=20)
    n_tr = R.state.y["train"].nnz
    nb = -(-R.n_rows // bs)
    t_batch = n_tr / nb
    bytes_dec = t_batch * (4 * 256 + 16) + min(bs, R.n_rows) * 256 * 4 * 2
    hbm, _ = measured_peaks()
    metrics = R.evaluate("test")
    # the decoder's tensor-core form on the same batch and plan (measured per BASELINE shape)
    tc_vs = None
    if eng.target.sorted:
        try:
            eng.set_decoder("tc")
            prof_tc = eng.h.profile_step(b=0, reps=10)
            eng.set_decoder("gather")
            keys = ("decoder_loss_dz3", "dw4_segments")
            tc_vs = {"gather_ms": {k: prof[k] for k in keys}, "tc_3xTF32_ms": {k: prof_tc[k] for k in keys},
                     "tc_over_gather": sum(prof_tc[k] for k in keys) / sum(prof[k] for k in keys),
                     "target_density": n_tr / (R.n_rows * R.state.n_cols)}
        except Exception as e:
            tc_vs = {"error": "{}: {}".format(type(e).__name__, e)}
    n_params = eng.h.n_params
    d_batch = float(eng.d_len.sum()) / nb
    by_class = step_algorithmic_bytes(t_batch, d_batch, min(R.state.n_cols, int(t_batch)), min(eng.n_enc, int(d_batch)),
                                      min(bs, R.n_rows), n_params)
    classes = {k: {"ms_per_launch": v, "algorithmic_bytes_per_launch": by_class[k],
                   "achieved_GBps": by_class[k] / (v * 1e-3) / 1e9,
                   "frac_of_hbm_peak": by_class[k] / (v * 1e-3) / 1e9 / hbm} for k, v in prof.items() if k in by_class}
    agg_bytes = E.bytes_per_round(R.K, n_tr, R.state.y["test"].nnz, R.n_rows, [len(c) for c in R.cols], R.state.n_cols,
                                  local_epochs, bs)
    out = {"control_name": control,
           "workload": "{}-shape (synthetic, SURVEY.md 8d): {} x {}, {} train / {} test entries, {} organizations, "
                       "batch {} rows, {} local epochs".format(data_name, R.n_rows, R.state.n_cols, n_tr,
                                                               R.state.y["test"].nnz, R.K, bs, local_epochs),
           "ms_per_round": ms, "value": visits / (ms / 1e3), "unit": UNIT,
           "host_ms_per_round": {k: 1e3 * v / n_rounds for k, v in phase.items()},
           "step_mode": eng.h.step_mode(), "decoder": eng.decoder, "step_kernel_ms": prof,
           "step_sum_us": 1e3 * sum(prof.values()),
           "decoder_roofline": {"algorithmic_bytes_per_launch": bytes_dec, "ms_per_launch": prof["decoder_loss_dz3"],
                                "achieved_GBps": bytes_dec / (prof["decoder_loss_dz3"] * 1e-3) / 1e9,
                                "frac_of_hbm_peak": bytes_dec / (prof["decoder_loss_dz3"] * 1e-3) / 1e9 / hbm},
           "step_classes": classes,
           "round_aggregate": {"algorithmic_bytes_per_round": agg_bytes, "achieved_GBps": agg_bytes / (ms * 1e-3) / 1e9,
                               "frac_of_hbm_peak": agg_bytes / (ms * 1e-3) / 1e9 / hbm},
           "tc_vs_gather": tc_vs,
           "test_metrics_after_{}_rounds".format(n_warm + n_rounds): metrics}
    R.close()
    del R
    torch.cuda.empty_cache()
    if cpu_leg:
        ref = reference_leg(os.cpu_count() or 1, 2, 1, budget_s=60.0, control=control, data=data_name)
        out["cpu_baseline"] = ref if "unavailable" in ref else {
            "value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": "reference", "round_s": ref["round_s"],
            "sample": ref["sample"]}
    return out


def run_e2e(args, data, rank, world, dev, state_dicts=False, n_steps=None):
    """Rounds through the drop-in API (host scipy CSR in / out) on ALL ranks: with torch.distributed initialised the
    drop-in classes shard the organizations over the ranks themselves (Organization.train / predict do real work only
    for this rank's organizations, Assist.update all-gathers the prediction vectors over NCCL; dropin/assist.py), so
    this is the call sequence of the reference's driver launched under torchrun. Host wall clock around every round
    (device synchronised), max over ranks."""
    from dmtcdr_b200 import dist as D
    from dmtcdr_b200 import engine as E
    from dmtcdr_b200 import runner
    from dmtcdr_b200.config import cfg

    E.XFER["h2d"] = E.XFER["d2h"] = 0
    n_steps = n_steps or max(1, min(args.steps, 3))
    n_warm = 2  # round 1 captures/instantiates the epoch graphs; round 2 still grows allocator pools
    state = {}
    times = []

    def on_round(t):
        torch.cuda.synchronize()
        now = time.perf_counter()
        if t > n_warm:
            times.append(now - state["t"])
        if t == n_warm:
            E.XFER["h2d"] = E.XFER["d2h"] = 0
        state["t"] = now

    D.barrier()
    t0 = time.perf_counter()
    state["t"] = t0
    res = runner.run_assist_experiment(data, CONTROL, seed=0, local_epochs=args.local_epochs, rounds=n_warm + n_steps,
                                       rng="device", on_round=on_round, materialize_state_dicts=state_dicts)
    sec = D.max_over_ranks(sum(times) / len(times), dev)
    h2d = D.sum_over_ranks(float(E.XFER["h2d"]), dev)
    d2h = D.sum_over_ranks(float(E.XFER["d2h"]), dev)
    K = 18
    visits = K * (args.local_epochs * data.train.nnz + data.train.nnz + data.test.nnz)
    return {"value": visits / sec, "unit": UNIT, "ms_per_step": 1e3 * sec,
            "h2d_bytes_per_step": int(h2d / n_steps), "d2h_bytes_per_step": int(d2h / n_steps),
            "api": "Assist.make_dataset / Organization.train / Organization.predict / Assist.update + test metrics, "
                   "host scipy CSR in and out; organizations sharded over the ranks inside the drop-in classes, NCCL "
                   "all-gather inside Assist.update; bytes summed over ranks",
            "n_gpus_used": world, "rng": "device (device-drawn init/dropout: the production mode, not the parity mode)",
            "state_dicts": "read back every round (the reference's per-round CPU copy, src/organization.py:177)"
                           if state_dicts else "lazy: left on the device until read (the reference copies 18 x 4.3 MB "
                                               "to the host every round; see e2e_with_state_dicts)",
            "rmse_last_round": res["metrics"][n_warm + n_steps].get("test/RMSE")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--local-epochs", dest="local_epochs", type=int, default=20)
    ap.add_argument("--configs", default="all", help="extra BASELINE configs as blocks of the line: all|none|douban|amazon")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
