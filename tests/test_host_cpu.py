"""CPU-side checks (run with -m "not gpu"): the C-ABI library loads and exports every declared symbol, the
reference-vocabulary config, synthetic data, epoch layouts, and the organization sharding / exchange on 2 gloo ranks."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import build, native

    path = build.build_native()
    lib = native.load()  # ctypes declares every entry point: a missing one raises AttributeError here
    with open(os.path.join(ROOT, "include", "dmt_b200.h")) as f:
        declared = set(re.findall(r"\b(dmt_[a-z0-9_]+)\s*\(", f.read()))
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dmt_[a-z0-9_]+)", out))
    assert declared and declared <= exported, sorted(declared - exported)
    assert lib.dmt_version() >= 100
    # only the CUDA runtime/driver and libc: no torch, no python in the boundary library
    ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd


def test_sass_is_sm100a_only():
    from dmtcdr_b200 import build

    out = subprocess.run(["cuobjdump", "-lelf", build.build_native()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_path():
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import native

    with pytest.raises(native.NativeError):
        native.ptr(torch.zeros(3))


def test_make_cfg_matches_reference_tables():
    from dmtcdr_b200.config import make_cfg

    c = make_cfg("ML1M_user_explicit_ae_0_genre_assist_constant-0.1_constant", install=False)
    assert c["num_organizations"] == 18 and c["local"]["batch_size"]["train"] == 500
    assert c["local"]["num_epochs"] == 20 and c["global"]["num_epochs"] == 10
    assert c["assist"]["ar_mode"] == "constant" and c["assist"]["ar"] == 0.1 and "cs" not in c and "pl" not in c
    c = make_cfg("Amazon_user_implicit_ae_0_genre_assist_constant-0.1_optim_0.5_dp-10", install=False)
    assert c["num_organizations"] == 4 and c["assist"]["aw_mode"] == "optim" and c["assist"]["match_rate"] == 0.5
    assert c["pl_mode"] == "dp" and c["pl_param"] == 10.0 and "cs" not in c
    c = make_cfg("Douban_item_explicit_ae_0_random-8_assist_optim-0.3_constant", install=False)
    assert c["num_organizations"] == 8 and c["local"]["batch_size"]["train"] == 1000


def test_synthetic_data_shape_and_split():
    from dmtcdr_b200 import synth

    d = synth.make_rating_data("tiny-ML100K", seed=0)
    M, N, nnz, G, P = synth.SHAPES["tiny-ML100K"]
    assert d.train.shape == (M, N) and d.train.nnz + d.test.nnz == nnz and d.train.nnz == int(nnz * 0.9)
    assert d.train.multiply(d.test).nnz == 0  # disjoint (user, item) pairs
    assert d.item_attr.shape == (N, G) and d.user_profile.shape == (M, P)
    (trd, trt), (ted, tet) = d.split("implicit")
    assert set(np.unique(trt.data)) <= {0.0, 1.0} and ted is trd
    d2 = synth.make_rating_data("tiny-ML100K", seed=0)
    assert (d.train != d2.train).nnz == 0


def test_epoch_layouts():
    from dmtcdr_b200 import engine as E

    rng = np.random.default_rng(0)
    d_len = rng.integers(0, 3, size=50)
    t_len = rng.integers(0, 4, size=50)
    d_len[[3, 4]] = 0
    t_len[[4]] = 0  # row 4 has nothing: dropped; row 3 has targets only
    perm = rng.permutation(50)
    batches = [perm[s:s + 16] for s in range(0, 50, 16)]
    a = E.EpochLayout(batches, d_len, t_len)
    b = E.FastEpochLayout(perm, 16, d_len, t_len)
    assert 4 not in a.rows and 4 not in b.rows and 3 in a.rows
    assert a.n_t == b.n_t == int(t_len.sum()) and a.n_d == b.n_d == int(d_len.sum())
    assert list(a.row_off) == list(b.row_off)
    for k in range(len(batches)):
        ra = a.rows[a.row_off[k]:a.row_off[k + 1]]
        rb = b.rows[b.row_off[k]:b.row_off[k + 1]]
        assert list(ra) == sorted(ra) and sorted(rb) == list(ra)
    assert a.active == b.active and a.d_per_batch == b.d_per_batch


def test_index_batches_consume_rng_like_a_dataloader():
    from dmtcdr_b200 import engine as E

    torch.manual_seed(3)
    b1 = E.index_batches(23, 5, True)
    after = torch.rand(1)
    torch.manual_seed(3)
    torch.empty((), dtype=torch.int64).random_()  # base seed of the loader iterator
    seed = int(torch.empty((), dtype=torch.int64).random_().item())  # RandomSampler's seed
    g = torch.Generator()
    g.manual_seed(seed)
    perm = torch.randperm(23, generator=g).numpy()
    assert np.array_equal(np.concatenate(b1), perm) and float(after) == float(torch.rand(1))


def test_org_blocks_cover_all_organizations():
    from dmtcdr_b200 import dist as D

    for K, world in [(18, 1), (18, 2), (18, 4), (18, 8), (3, 8), (64, 8)]:
        seen = []
        for r in range(world):
            orgs, c = D.org_block(K, world, r)
            assert len(orgs) <= c
            seen += orgs
        assert seen == list(range(K))


def test_balanced_assignment_leaves_no_rank_idle():
    """dist.assign_orgs: every organization owned once, counts differ by at most one (18 organizations on 8 ranks:
    3,3,2,2,2,2,2,2 instead of the contiguous blocks' 3,3,3,3,3,3,0,0), heavier organizations spread first, and every
    organization's row of the rank-blocked O_full lies inside its owner's block."""
    from dmtcdr_b200 import dist as D

    rng = np.random.default_rng(0)
    for K, world in [(18, 1), (18, 2), (18, 4), (18, 8), (3, 8), (64, 8), (4, 3)]:
        costs = rng.uniform(1, 10, K)
        mine, chunk, org_row = D.assign_orgs(costs, world)
        assert sorted(sum(mine, [])) == list(range(K))
        counts = [len(m) for m in mine]
        assert max(counts) - min(counts) <= 1 and chunk == max(counts)
        assert len(set(org_row)) == K
        for r, m in enumerate(mine):
            assert m == sorted(m)
            for j, k in enumerate(m):
                assert org_row[k] == r * chunk + j
        if K >= 2 * world:
            loads = [sum(costs[k] for k in m) for m in mine]
            assert max(loads) <= 1.5 * (sum(loads) / world)
    assert D.assign_orgs([1.0] * 18, 8)[0] == D.assign_orgs([1.0] * 18, 8)[0]  # deterministic


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import dmtcdr_b200
from dmtcdr_b200 import dist as D
rank, world, _ = D.init_from_env(backend="gloo")
K, nnz = 5, 37
mine, chunk, org_row = D.assign_orgs([3.0, 1.0, 4.0, 1.0, 5.0], world)
orgs = mine[rank]
O = {{k: torch.zeros(chunk * world, nnz) for k in ("train", "test")}}
for k, O_k in O.items():
    for o in orgs:
        O_k[org_row[o]] = torch.arange(nnz, dtype=torch.float32) + 100 * o + (1000 if k == "test" else 0)
D.exchange_outputs(O, chunk, rank, world)
for k, O_k in O.items():
    for o in range(K):
        want = torch.arange(nnz, dtype=torch.float32) + 100 * o + (1000 if k == "test" else 0)
        assert torch.equal(O_k[org_row[o]], want), (rank, k, o)
    pad = [r for r in range(chunk * world) if r not in org_row]
    assert float(O_k[pad].abs().sum()) == 0.0
assert D.max_over_ranks(float(rank), "cpu") == float(world - 1)
D.barrier()
sys.stdout.write("rank%dok\n" % rank)
sys.stdout.flush()
"""


def test_exchange_on_two_gloo_ranks(tmp_path):
    """N>1 path on CPU: the in-place all-gather of organization outputs, world_size 2, gloo."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


def test_distribute_slices_the_split_entity():
    """models.distribute (reference src/models/utils.py:17-40): joint tables are row-sliced into the local models."""
    import dmtcdr_b200
    from dmtcdr_b200.config import make_cfg

    cfg = make_cfg("ML100K_user_explicit_mf_0_genre_joint", device="cpu")
    cfg["info_size"] = None
    models, _, _ = dmtcdr_b200.use_dropin()
    torch.manual_seed(0)
    joint = models.mf(10, 7)
    split = [torch.tensor([0, 3, 4]), torch.tensor([1, 2, 5, 6])]
    local = [models.mf(10, len(s)) for s in split]
    models.distribute(joint, local, split)
    for s, m in zip(split, local):
        assert torch.equal(m.item_weight.weight, joint.item_weight.weight[s])
        assert torch.equal(m.item_bias.weight, joint.item_bias.weight[s])
        assert torch.equal(m.user_weight.weight, joint.user_weight.weight)
        assert torch.equal(m.bias, joint.bias)


def test_lazy_csr_and_pending_flush():
    """Drop-in predict() outputs are scipy CSRs whose values arrive lazily; deferred log callbacks run at flush."""
    import pickle

    import numpy as np
    from scipy.sparse import csr_matrix

    import dmtcdr_b200

    _, org_mod, _ = dmtcdr_b200.use_dropin()
    m = org_mod.LazyCSR((np.empty(4, np.float32), np.array([0, 1, 0, 2], np.int32), np.array([0, 2, 4], np.int32)),
                        shape=(2, 3), copy=False)
    dest = m.__dict__['_dmt_data']
    calls = []

    def force():
        calls.append(1)
        np.copyto(dest, np.array([1, 2, 3, 4], np.float32))

    m.__dict__['_dmt_force'] = force
    assert m.nnz == 4 and m.shape == (2, 3) and not calls        # structure queries do not wait for the values
    assert isinstance(m, csr_matrix)
    assert m.toarray().tolist() == [[1.0, 2.0, 0.0], [3.0, 0.0, 4.0]] and calls == [1]
    assert m.data.tolist() == [1.0, 2.0, 3.0, 4.0] and calls == [1]  # forced once
    p = pickle.loads(pickle.dumps(m))
    assert type(p) is csr_matrix and p.data.tolist() == [1.0, 2.0, 3.0, 4.0]
    seen = []
    org_mod._PENDING.append(lambda: seen.append('a'))
    org_mod._PENDING.append(lambda: seen.append('b'))
    org_mod.flush_pending()
    org_mod.flush_pending()
    assert seen == ['a', 'b']


def test_whole_round_layout_equals_concatenated_epochs():
    """FastEpochLayout(epoch_len=...) — one plan for all local epochs — lists exactly the per-epoch layouts back to back."""
    import numpy as np

    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import engine as E

    rng = np.random.default_rng(0)
    n, bs, ep = 103, 10, 4
    d_len, t_len = rng.integers(0, 3, n), rng.integers(0, 4, n)
    perms = [rng.permutation(n) for _ in range(ep)]
    lays = [E.FastEpochLayout(p, bs, d_len, t_len) for p in perms]
    whole = E.FastEpochLayout(np.concatenate(perms), bs, d_len, t_len, epoch_len=n)
    assert np.array_equal(whole.rows, np.concatenate([l.rows for l in lays]))
    base = np.cumsum([0] + [len(l.rows) for l in lays[:-1]])
    glob = np.concatenate([l.row_off[:-1] + b for l, b in zip(lays, base)] + [np.array([sum(len(l.rows) for l in lays)])])
    assert np.array_equal(whole.row_off, glob)
    assert whole.active == sum([l.active for l in lays], [])
    assert whole.d_per_batch == sum([l.d_per_batch for l in lays], [])
    assert whole.n_t == sum(l.n_t for l in lays) and whole.n_d == sum(l.n_d for l in lays)


def test_ctypes_signatures_match_the_header():
    """Every prototype in include/dmt_b200.h has a ctypes declaration in native.py with the same number of arguments
    (an arity mismatch would only show up as a crash on the GPU box)."""
    import re

    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import native

    lib = native.load()
    with open(os.path.join(ROOT, "include", "dmt_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    protos = re.findall(r"\b(dmt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) > 40
    checked = 0
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue  # exported but not bound from Python
        assert len(fn.argtypes) == n, "{}: header has {} arguments, native.py declares {}".format(name, n,
                                                                                                   len(fn.argtypes))
        checked += 1
    assert checked > 40


def test_bench_roofline_inputs_parse():
    """bench.py's roofline block reads the committed ncu raw page (profiles/r2_ncu_raw_fused.csv) at run time: every
    kernel of the fused step must be found in it with non-zero DRAM / L2->SM bytes, and the per-class algorithmic bytes
    follow SURVEY.md section 8d (1040 B per decoder target, 28 + 4 B per Adam parameter)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per = bench.ncu_per_kernel()
    for cls, names in bench.CLASS_KERNEL.items():
        kern = names[1]  # the default (register-load) form
        assert kern in per, (cls, kern, sorted(per))
        assert per[kern]["launches"] >= 1 and per[kern]["us"] > 0 and per[kern]["l2_to_sm_bytes"] > 0
    b = bench.step_algorithmic_bytes(t_batch=1000.0, d_batch=100.0, n_seg_t=50, n_seg_d=10, B=500, n_params=10)
    assert b["decoder_loss_dz3"] == 1000 * (4 * 256 + 16) + 4 * 500 * 256 * 2
    assert b["clip_adam"] == 32 * 10 and b["grad_norm"] == 40
    with pytest.raises(FileNotFoundError):
        bench.ncu_per_kernel(os.path.join(ROOT, "profiles", "does_not_exist.csv"))


def test_round_layout_equals_per_epoch_layouts():
    """engine.RoundLayout (one pass per organization and round) slices into exactly the per-epoch FastEpochLayouts."""
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import engine as E

    rng = np.random.default_rng(4)
    for n_rows, bs in ((1037, 100), (500, 500), (6040, 500), (7, 3)):
        d_len = rng.integers(0, 3, n_rows).astype(np.int64)
        t_len = rng.integers(0, 4, n_rows).astype(np.int64)
        n_ep = 5
        perms = [rng.permutation(n_rows) for _ in range(n_ep)]
        R = E.RoundLayout(np.concatenate(perms), bs, d_len, t_len, n_rows, n_ep)
        per = [E.FastEpochLayout(p, bs, d_len, t_len) for p in perms]
        nb = R.nb_epoch
        assert R.n_batches == sum(len(l.active) for l in per)
        assert (R.n_t_total, R.n_d_total) == (sum(l.n_t for l in per), sum(l.n_d for l in per))
        for e, l in enumerate(per):
            assert np.array_equal(R.rows[R.row_edges[e]:R.row_edges[e + 1]], l.rows)
            assert np.array_equal(R.off_local[e], l.row_off) and len(l.row_off) == nb + 1
            assert (R.n_t[e], R.n_d[e]) == (l.n_t, l.n_d)
            assert R.whole.active[e * nb:(e + 1) * nb] == l.active
        glob = np.concatenate([[0], np.cumsum([c for l in per for c in l.batch_rows])])
        assert np.array_equal(R.off_global, glob)


def test_dropin_round_layout_prefetch_is_deterministic():
    """The drop-in prepares the NEXT round's batch layout on a background thread (dropin/organization.py:
    _prefetch_round_layout): prefetched and foreground layouts are the same object-for-object, and both equal the
    permutations the per-(seed, organization, round) host generator draws."""
    import dmtcdr_b200
    from dmtcdr_b200 import engine as E
    from dmtcdr_b200.config import make_cfg

    make_cfg("ML100K_user_explicit_ae_0_genre_assist_constant-0.1_constant", device="cpu", seed=3)
    dmtcdr_b200.use_dropin()
    import organization as org_mod

    rng = np.random.default_rng(0)
    n_own, n_ep, bs = 333, 4, 50
    d_len = rng.integers(0, 3, n_own).astype(np.int64)
    t_len = rng.integers(0, 5, n_own).astype(np.int64)
    org_mod._prefetch_round_layout(2, 5, n_own, n_ep, bs, d_len, t_len)
    a = org_mod._round_layout(2, 5, n_own, n_ep, bs, d_len, t_len)      # takes the prefetched future
    b = org_mod._round_layout(2, 5, n_own, n_ep, bs, d_len, t_len)      # builds it in the foreground
    g = torch.Generator()
    g.manual_seed(E.he_seed(3, 2, 5, 1 << 21) & (2 ** 63 - 1))
    perms = np.concatenate([torch.randperm(n_own, generator=g).numpy() for _ in range(n_ep)])
    c = E.FastEpochLayout(perms, bs, d_len, t_len, epoch_len=n_own)
    for x in (a, b):
        assert np.array_equal(x.rows, c.rows) and np.array_equal(x.row_off, c.row_off)
        assert x.active == c.active and (x.n_t, x.n_d) == (c.n_t, c.n_d)
    other = org_mod._round_layout(2, 6, n_own, n_ep, bs, d_len, t_len)
    assert not np.array_equal(other.rows, c.rows)


def test_reference_arm_harness_runs_the_unmodified_reference(tmp_path):
    """baseline/ref_arm.py (the bench's `--impl reference` and `cpu_baseline` leg) imports the UNMODIFIED reference
    (baseline/_ref/src, vendored by __graft_entry__.build(), or /root/reference/src in the build container) and times
    its own Assist.make_dataset / Organization.train / predict / Assist.update on the host: here on the tiny Douban-shape
    data, one step. Skipped only where no copy of the reference exists."""
    import json as _json

    if not any(os.path.exists(os.path.join(p, "organization.py"))
               for p in (os.path.join(ROOT, "baseline", "_ref", "src"), "/root/reference/src")):
        pytest.skip("no copy of the reference sources")
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_arm.py"), "--control",
           "Douban_user_explicit_ae_0_genre_assist_constant-0.3_constant", "--data", "tiny-Douban", "--device", "cpu",
           "--threads", "2", "--steps", "1", "--warmup", "0", "--budget-s", "30"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    lines = [l for l in out.stdout.splitlines() if l.startswith("REF_ARM_JSON ")]
    assert out.returncode == 0 and lines, (out.stdout[-500:], out.stderr[-1500:])
    res = _json.loads(lines[-1][len("REF_ARM_JSON "):])
    assert res["kind"] == "reference" and res["cores"] == 2 and res["steps_done"] == 1
    assert res["value"] > 0 and res["round_s"] > 0
    for k in ("t_make_dataset_s", "t_train_epoch_s", "t_predict_s", "t_update_s"):
        assert res[k] > 0, k
