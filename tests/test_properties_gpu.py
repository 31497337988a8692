"""Size-independent properties at BASELINE.json's full ML1M size (6040 x 3706, ~1 M ratings, K = 18), where the
oracle would take minutes: sortedness / permutation validity / idempotence of the segment sort, linearity and identities
of the coordinator kernels, and a checksum-of-checksums of the organization-major combination."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dmtcdr_b200  # noqa: F401
    from dmtcdr_b200 import native, synth

    native.load()
    data = synth.make_rating_data("ML1M", seed=0)
    return native, data


def test_sort_segments_full_size(env):
    nat, data = env
    tr = data.train
    keys = torch.from_numpy(tr.indices.astype(np.int32)).cuda()  # 900 188 item ids in row-major order
    perm, seg_key, seg_off, n_seg = nat.sort_segments(keys, tr.shape[1])
    n, ns = keys.numel(), int(n_seg)
    p = perm[:n].long()
    sk = keys[p]
    assert bool((sk[1:] >= sk[:-1]).all())  # sorted
    assert torch.equal(torch.sort(p).values, torch.arange(n, device="cuda"))  # a permutation
    same = sk[1:] == sk[:-1]
    assert bool((p[1:][same] > p[:-1][same]).all())  # stable inside every segment
    assert ns == int(torch.unique(keys).numel()) and int(seg_off[ns]) == n
    counts = torch.bincount(keys.long(), minlength=tr.shape[1])
    assert torch.equal((seg_off[1:ns + 1] - seg_off[:ns]).long(), counts[seg_key[:ns].long()])
    # idempotence: sorting the sorted keys is the identity permutation
    perm2, *_ = nat.sort_segments(sk.contiguous(), tr.shape[1])
    assert torch.equal(perm2[:n].long(), torch.arange(n, device="cuda"))


def test_residual_identities_full_size(env):
    nat, data = env
    y = torch.from_numpy(data.train.data).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    F = torch.randn(y.numel(), device="cuda", generator=g) + 3
    r = nat.residual(F, y, 0)
    assert torch.equal(r, 2 * (y - F))  # explicit: exactly 2(y-F) in fp32
    assert float(nat.residual(y, y, 0).abs().max()) == 0.0
    rc = nat.residual(F, y, 0, 1.0)
    assert torch.equal(rc, r.clamp(-1, 1))
    yb = (y >= 3.5).float()
    ri = nat.residual(F, yb, 1)
    assert float((ri - (yb - torch.sigmoid(F))).abs().max()) < 2e-7
    assert float(ri.abs().max()) <= 1.0


def test_combine_linearity_and_checksum_full_size(env):
    nat, data = env
    tr = data.train
    K, n_cols, nnz = 18, tr.shape[1], tr.nnz
    rng = np.random.default_rng(1)
    owner = torch.from_numpy(rng.integers(0, K, size=n_cols).astype(np.int32)).cuda()
    col = torch.from_numpy(tr.indices.astype(np.int32)).cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    F0 = torch.randn(nnz, device="cuda", generator=g)
    O = torch.randn(K, nnz, device="cuda", generator=g)
    rate = torch.rand(n_cols, device="cuda", generator=g)
    S = torch.softmax(torch.randn(K, K, device="cuda", generator=g), -1).contiguous()
    F1 = nat.assist_combine(F0, O, col, owner, rate, S)
    # linear in O: combine(F0, a*O) - F0 == a * (combine(F0, O) - F0)
    F2 = nat.assist_combine(F0, (2 * O).contiguous(), col, owner, rate, S)
    assert float(((F2 - F0) - 2 * (F1 - F0)).abs().max()) < 1e-5
    # zero rate is the identity; identical rows of O collapse the softmax
    assert torch.equal(nat.assist_combine(F0, O, col, owner, torch.zeros_like(rate), S), F0)
    same = O[:1].repeat(K, 1).contiguous()
    F3 = nat.assist_combine(F0, same, col, owner, rate, S)
    assert float((F3 - (F0 + rate[col.long()] * O[0])).abs().max()) < 1e-5
    # checksum of checksums: sum_p (F1-F0)[p] == sum_i sum_j S[i,j] * sum_{p owned by i} rate*O[j,p]  (float64 on the host)
    own_p = owner[col.long()].long()
    w = (rate[col.long()].double() * O.double())  # [K, nnz]
    per_owner = torch.zeros(K, K, dtype=torch.float64, device="cuda").index_add_(1, own_p, w)  # [j, i]
    want = float((per_owner.t() * S.double()).sum())
    got = float((F1.double() - F0.double()).sum())
    assert abs(got - want) <= 1e-6 * max(1.0, abs(want)) + 1e-3
