"""Layout diagnostics of the tcgen05 building blocks (scripts only; not part of the product or the tests).

Feeds structured integer inputs (exact in TF32) through dmt_dense_fwd_tc so a wrong swizzle / descriptor / TMEM lane
mapping shows up as a recognisable permutation instead of just "wrong numbers".
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dmtcdr_b200  # noqa: F401
from dmtcdr_b200 import native as nat

torch.manual_seed(0)
dev = "cuda"


def report(name, got, ref):
    err = (got.double() - ref.double()).abs().max().item()
    scale = ref.double().abs().max().item()
    print("{:40s} max|err| {:.3e}  (max|ref| {:.3e})".format(name, err, scale), flush=True)
    return err <= 1e-4 * max(scale, 1e-30)


def identity_probe(M, N, K, passes):
    X = (torch.arange(M).view(M, 1) % 16 * 32 + torch.arange(K).view(1, K) % 32 + 1).float()
    W = torch.zeros(N, K)
    for n in range(min(N, K)):
        W[n, n] = 1.0
    Y, _ = nat.dense_fwd_tc(X.to(dev), W.to(dev), None, 0, passes=passes)
    torch.cuda.synchronize()
    ref = X @ W.t()
    ok = report("identity M{} N{} K{} p{}".format(M, N, K, passes), Y.cpu(), ref)
    if not ok:
        Yc = Y.cpu()
        print("got[0:4, 0:12]\n", Yc[0:4, 0:12])
        print("ref[0:4, 0:12]\n", ref[0:4, 0:12])
        bad = (Yc != ref).nonzero()
        print("first mismatches (row, col):", bad[:16].tolist(), "count", len(bad))
        for m, n in bad[:8].tolist():
            src = (X[m] == Yc[m, n]).nonzero().flatten().tolist()
            print("  Y[{},{}] = {} (ref {}), equals X[{}, k] for k in {}".format(m, n, Yc[m, n].item(), ref[m, n].item(), m, src))
    return ok


def random_probe(M, N, K, passes):
    X = torch.randn(M, K)
    W = torch.randn(N, K)
    Y, _ = nat.dense_fwd_tc(X.to(dev), W.to(dev), None, 0, passes=passes)
    torch.cuda.synchronize()
    ref = X.double() @ W.double().t()
    err = (Y.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    print("random M{} N{} K{} p{}: rel err {:.3e}".format(M, N, K, passes, err), flush=True)


if __name__ == "__main__":
    assert nat.load().dmt_check_device() == 0
    ok = identity_probe(128, 128, 32, 1)
    ok = identity_probe(128, 128, 32, 3) and ok
    ok = identity_probe(128, 128, 64, 1) and ok
    ok = identity_probe(256, 256, 32, 1) and ok
    for p in (1, 3):
        random_probe(128, 128, 32, p)
        random_probe(500, 256, 128, p)
    # transposed staging: dW = dY^T X
    dY = torch.randn(96, 128)
    X = torch.randn(96, 64)
    for p in (1, 3):
        dW, db = nat.dense_bwd_w_tc(dY.to(dev), X.to(dev), passes=p)
        torch.cuda.synchronize()
        ref = dY.double().t() @ X.double()
        print("bwd_w p{}: rel err {:.3e}".format(p, (dW.cpu().double() - ref).abs().max().item() / ref.abs().max().item()),
              flush=True)
    print("debug_tc done, identity ok =", ok)
