"""ctypes binding of libdmt_b200.so — the C-ABI declared in include/dmt_b200.h.

There is NO fallback: if the shared library is missing or a call fails, an exception is raised. torch is used
only as the owner of device memory and streams (``tensor.data_ptr()``, ``current_stream().cuda_stream``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

_LIB = None

c_i32p = C.c_void_p  # device pointers are passed as opaque addresses
c_f32p = C.c_void_p


class NativeError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """Load the shared library (building it in-tree with nvcc when absent). Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if not build_if_missing:
            raise NativeError("libdmt_b200.so is missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
        _build.build_native()
    lib = C.CDLL(path)
    _declare(lib)
    _LIB = lib
    return lib


def _declare(lib):
    P, I, L, F, D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    lib.dmt_last_error.restype = C.c_char_p
    lib.dmt_last_error.argtypes = []
    sig = {
        "dmt_version": (I, []),
        "dmt_check_device": (I, []),
        "dmt_residual": (I, [P, P, P, L, I, F, P]),
        "dmt_assist_combine": (I, [P, P, P, P, P, P, P, P, L, I, P, P, P]),
        "dmt_assist_gather_view": (I, [P, P, P, P, P, L, L, I, I, L, P, P, P, P, P]),
        "dmt_assist_rows_fwd": (I, [P, L, L, P, P, P, P, L, I, P, P, P]),
        "dmt_assist_rows_scratch_floats": (L, [I]),
        "dmt_assist_rows_bwd": (I, [P, L, L, P, P, P, P, P, P, P, P, P, L, I, I, P, P, P, P]),
        "dmt_assist_scratch_floats": (L, [I]),
        "dmt_assist_loss_grad": (I, [P, P, P, P, P, P, L, I, I, I, P, P, P, P, P]),
        "dmt_assist_fit_work_floats": (L, [I, I, I]),
        "dmt_assist_fit": (I, [P, P, P, P, L, I, I, I, P, I, I, F, I, I, I, P, P, P]),
        "dmt_base_fit": (I, [P, P, L, P, P, P]),
        "dmt_base_predict": (I, [P, P, C.c_int32, P, L, I, F, P, P, P]),
        "dmt_sort_segments_temp_bytes": (L, [L]),
        "dmt_sort_segments": (I, [P, L, C.c_int32, P, P, P, P, P, L, P]),
        "dmt_segment_reduce_rows": (I, [P, P, P, P, L, P, P, P, I, P, P, P]),
        "dmt_sqnorm_scratch_floats": (L, []),
        "dmt_sqnorm": (I, [P, L, P, P, P]),
        "dmt_adam_clip_step": (I, [P, P, P, P, L, P, F, D, D, D, D, D, L, P, P]),
        "dmt_mf_scratch_floats": (L, []),
        "dmt_mf_fwd": (I, [P, P, P, L, P, P, P, P, P, P, P, P, P, I, I, P, P, P, P, P, P]),
        "dmt_mf_bwd_table": (I, [P, P, P, P, I, P, F, P, P, P, P, L, P, P, P, P]),
        "dmt_mf_bwd_side": (I, [P, L, P, P, I, P, F, P, P, P]),
        "dmt_embed_fwd": (I, [P, L, P, P, I, P, I, I, P]),
        "dmt_embed_bwd": (I, [P, I, I, I, P, P, P, P, L, P, P, P]),
        "dmt_weighted_colsum_scratch_floats": (L, [I]),
        "dmt_weighted_colsum": (I, [P, F, P, L, I, I, P, P, P]),
        "dmt_loss_fwd": (I, [P, P, L, I, P, P, P, P]),
        "dmt_dense_fwd": (I, [P, P, P, P, P, P, F, I, I, I, I, P]),
        "dmt_dense_bwd_x": (I, [P, P, P, P, F, P, I, I, I, I, P]),
        "dmt_dense_bwd_w": (I, [P, P, P, P, I, I, I, P]),
        "dmt_dense_fwd_tc": (I, [P, P, P, P, P, P, F, I, I, I, I, I, P]),
        "dmt_dense_bwd_x_tc": (I, [P, P, P, P, F, P, I, I, I, I, I, P]),
        "dmt_dense_bwd_w_tc": (I, [P, P, P, P, I, I, I, I, P]),
        "dmt_ae_decoder_tc_scratch_floats": (L, [I, I, I]),
        "dmt_ae_decoder_tc": (I, [P, I, P, P, P, P, P, P, I, I, I, P, I, P, P, P, P, P, P, I, P, P]),
        "dmt_org_set_decoder_mode": (I, [P, I, I]),
        "dmt_org_set_fanout": (I, [P, I]),
        "dmt_org_set_decoder_blocks": (I, [P, I]),
        "dmt_org_set_gather_mode": (I, [P, I]),
        "dmt_org_set_pdl": (I, [P, I]),
        "dmt_org_set_row_tile": (I, [P, I]),
        "dmt_org_gather_mode": (I, [P]),
        "dmt_org_set_step_mode": (I, [P, I]),
        "dmt_org_step_mode": (I, [P]),
        "dmt_ae_encoder_fwd": (I, [P, I, P, P, P, P, P, I, P, P]),
        "dmt_ae_decoder_fwd": (I, [P, I, P, P, P, P, P, P, I, I, P, P, P, P, P, I, P]),
        "dmt_eval_blocks": (I, [P, P, P, I, I, I, I, P, P, P]),
        "dmt_privacy_temp_bytes": (L, [L]),
        "dmt_privacy": (I, [P, L, I, F, C.c_uint64, P, P, P, L, P]),
        "dmt_org_create": (I, [C.POINTER(P), I, I, I, I, I, P, P, P, L, P, P, L, I, I, I, P]),
        "dmt_group_create": (I, [C.POINTER(P), C.POINTER(P), I, P]),
        "dmt_group_destroy": (I, [P]),
        "dmt_group_train": (I, [P, C.POINTER(P), C.POINTER(P), I, I, C.POINTER(L), C.POINTER(L), C.POINTER(C.c_uint64),
                                D, D, D, D, D, F, C.POINTER(P)]),
        "dmt_group_sync": (I, [P]),
        "dmt_group_stream": (P, [P]),
        "dmt_group_wait_stream": (I, [P, P]),
        "dmt_org_destroy": (I, [P]),
        "dmt_org_num_params": (L, [P]),
        "dmt_org_set_params": (I, [P, P]),
        "dmt_org_get_params": (I, [P, P]),
        "dmt_org_set_target": (I, [P, P]),
        "dmt_org_train_epoch": (I, [P, P, P, I, I, L, L, P, C.c_uint64, D, D, D, D, D, F, P]),
        "dmt_org_predict": (I, [P, P, P, P, P, P, I, P]),
        "dmt_org_sync": (I, [P]),
        "dmt_org_stream": (P, [P]),
        "dmt_org_wait_stream": (I, [P, P]),
        "dmt_org_signal_stream": (I, [P, P]),
        "dmt_launch_count": (L, []),
        "dmt_org_profile_classes": (I, []),
        "dmt_org_profile_step": (I, [P, I, I, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync: fail loudly
        fn.restype = res
        fn.argtypes = args


EXPORTS = None  # filled lazily for the symbol test


def check(rc, what):
    if rc != 0:
        msg = load().dmt_last_error().decode() if _LIB is not None else ""
        raise NativeError("{} failed (code {}): {}".format(what, rc, msg))


def ptr(t):
    """Device address of a tensor (None -> NULL). The tensor must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("dmtcdr_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise NativeError("non-contiguous tensor passed to a native kernel")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def i32(t):
    return t.to(torch.int32).contiguous()


def f32(t):
    return t.to(torch.float32).contiguous()


LOSS_KIND = {"explicit": 0, "implicit": 1}

# ------------------------------------------------------------------------------------------ thin wrappers


def residual(F, y, loss_kind, clamp=0.0, out=None):
    lib = load()
    out = torch.empty_like(F) if out is None else out
    check(lib.dmt_residual(ptr(F), ptr(y), ptr(out), F.numel(), loss_kind, float(clamp), stream()), "dmt_residual")
    return out


def assist_combine(F_old, O, col, owner, rate_col, S, match_end=None, out=None, org_row=None, S_cold=None, K=None):
    """O: [rows x nnz]; K organizations (default: rows) whose rows are org_row[j] (default j)."""
    lib = load()
    rows, nnz = O.shape
    K = rows if K is None else K
    out = torch.empty_like(F_old) if out is None else out
    check(lib.dmt_assist_combine(ptr(F_old), ptr(O), ptr(col), ptr(owner), ptr(rate_col), ptr(S), ptr(match_end),
                                 ptr(out), nnz, K, ptr(org_row), ptr(S_cold), stream()), "dmt_assist_combine")
    return out


def assist_gather_view(F_old, y, O, pos, rank, owner, n_match, org_row=None, K=None):
    lib = load()
    rows, nnz = O.shape
    K = rows if K is None else K
    n = pos.numel()
    h = torch.empty(n, device=O.device, dtype=torch.float32)
    t = torch.empty_like(h)
    V = torch.empty(K, n, device=O.device, dtype=torch.float32)
    check(lib.dmt_assist_gather_view(ptr(F_old), ptr(y), ptr(O), ptr(pos), ptr(rank), nnz, n, K, owner, n_match,
                                     ptr(h), ptr(t), ptr(V), ptr(org_row), stream()), "dmt_assist_gather_view")
    return h, t, V


def assist_rows_fwd(out, stride_e, stride_j, history, idx, rate, w, n, K, want_q=True):
    """models.Assist forward over n entries (cold-start aware) -> target, q."""
    lib = load()
    tgt = torch.empty(n, device=history.device, dtype=torch.float32)
    q = torch.empty(n, device=history.device, dtype=torch.float32) if want_q else None
    check(lib.dmt_assist_rows_fwd(ptr(out), stride_e, stride_j, ptr(history), ptr(idx), ptr(rate), ptr(w), n, K,
                                  ptr(tgt), ptr(q), stream()), "dmt_assist_rows_fwd")
    return tgt, q


def assist_rows_bwd(out, stride_e, stride_j, idx, rate, w, q, delta, seg, n, K, want_rate=True, want_w=True):
    """-> d_rate [n_rate] or None, d_w [K] or None. seg = sort_segments(idx, n_rate)."""
    lib = load()
    n_rate = rate.numel()
    dev = rate.device
    d_rate = torch.empty(n_rate, device=dev, dtype=torch.float32) if want_rate else None
    d_w = torch.empty(K, device=dev, dtype=torch.float32) if want_w else None
    scratch = torch.empty(lib.dmt_assist_rows_scratch_floats(K), device=dev, dtype=torch.float32)
    perm, seg_key, seg_off, n_seg = seg
    check(lib.dmt_assist_rows_bwd(ptr(out), stride_e, stride_j, ptr(idx), ptr(rate), ptr(w), ptr(q), ptr(delta),
                                  ptr(perm), ptr(seg_key), ptr(seg_off), ptr(n_seg), n, K, n_rate, ptr(d_rate),
                                  ptr(d_w), ptr(scratch), stream()), "dmt_assist_rows_bwd")
    return d_rate, d_w


def assist_fit(h, t, V, seg_off, params, ar_optim, aw_optim, loss_kind, lr=0.1, steps=10, max_iter=20, history=100,
               work=None, scratch=None):
    """Enqueue one owner's whole L-BFGS fit (dmt_assist_fit): params = [rate | w] on the device, updated in place; no
    host synchronisation. Returns (work, scratch) so that the caller keeps them alive until the stream has run."""
    lib = load()
    K, n = V.shape
    n_rate = params.numel() - K
    if scratch is None:
        scratch = torch.empty(lib.dmt_assist_scratch_floats(K), device=V.device, dtype=torch.float32)
    if work is None:
        work = torch.empty(lib.dmt_assist_fit_work_floats(n_rate, K, history), device=V.device, dtype=torch.float32)
    check(lib.dmt_assist_fit(ptr(h), ptr(t), ptr(V), ptr(seg_off), n, n_rate, K, loss_kind, ptr(params),
                             int(bool(ar_optim)), int(bool(aw_optim)), float(lr), int(steps), int(max_iter),
                             int(history), ptr(work), ptr(scratch), stream()), "dmt_assist_fit")
    return work, scratch


def assist_loss_grad(h, t, V, seg_off, rate, w, loss_kind, scratch=None):
    lib = load()
    K, n = V.shape
    n_rate = rate.numel()
    if scratch is None:
        scratch = torch.empty(lib.dmt_assist_scratch_floats(K), device=V.device, dtype=torch.float32)
    out = torch.empty(1 + n_rate + K, device=V.device, dtype=torch.float32)
    d_rate = out[1:1 + n_rate]
    d_w = out[1 + n_rate:]
    d_rate.zero_()
    check(lib.dmt_assist_loss_grad(ptr(h), ptr(t), ptr(V), ptr(seg_off), ptr(rate), ptr(w), n, n_rate, K, loss_kind,
                                   ptr(out), d_rate.data_ptr(), d_w.data_ptr(), ptr(scratch), stream()),
          "dmt_assist_loss_grad")
    return out  # [loss, d_rate..., d_w...]


def base_fit(idx, rating, base, count):
    check(load().dmt_base_fit(ptr(idx), ptr(rating), idx.numel(), ptr(base), ptr(count), stream()), "dmt_base_fit")


def base_predict(base, count, target_idx, implicit, implicit_count=0.0):
    out = torch.empty(target_idx.numel(), device=base.device, dtype=torch.float32)
    scratch = torch.empty(2, device=base.device, dtype=torch.float32)
    check(load().dmt_base_predict(ptr(base), ptr(count), base.numel(), ptr(target_idx), target_idx.numel(),
                                  int(implicit), float(implicit_count), ptr(out), ptr(scratch), stream()),
          "dmt_base_predict")
    return out


def sort_segments(keys, key_bound):
    """-> perm, seg_key, seg_off, n_seg (device int32 tensors; seg_* sized for the worst case)."""
    lib = load()
    n = keys.numel()
    dev = keys.device
    perm = torch.empty(max(n, 1), device=dev, dtype=torch.int32)
    seg_key = torch.empty(max(n, 1), device=dev, dtype=torch.int32)
    seg_off = torch.zeros(n + 2, device=dev, dtype=torch.int32)
    n_seg = torch.zeros(1, device=dev, dtype=torch.int32)
    nbytes = lib.dmt_sort_segments_temp_bytes(n)
    temp = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    check(lib.dmt_sort_segments(ptr(keys), n, int(key_bound), ptr(perm), ptr(seg_key), ptr(seg_off), ptr(n_seg),
                                ptr(temp), nbytes, stream()), "dmt_sort_segments")
    return perm, seg_key, seg_off, n_seg


def segment_reduce_rows(perm, seg_key, seg_off, n_seg, n_seg_max, coef, src_row, src, grad, bias_grad=None):
    width = src.shape[-1]
    check(load().dmt_segment_reduce_rows(ptr(perm), ptr(seg_key), ptr(seg_off), ptr(n_seg), n_seg_max, ptr(coef),
                                         ptr(src_row), ptr(src), width, ptr(grad), ptr(bias_grad), stream()),
          "dmt_segment_reduce_rows")


def sqnorm(g):
    lib = load()
    out = torch.empty(1, device=g.device, dtype=torch.float32)
    scratch = torch.empty(lib.dmt_sqnorm_scratch_floats(), device=g.device, dtype=torch.float32)
    check(lib.dmt_sqnorm(ptr(g), g.numel(), ptr(out), ptr(scratch), stream()), "dmt_sqnorm")
    return out


def adam_clip_step(w, g, m, v, step, sqnorm_t=None, max_norm=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                   weight_decay=5e-4):
    scratch = torch.empty(8, device=w.device, dtype=torch.float32)
    check(load().dmt_adam_clip_step(ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), ptr(sqnorm_t), float(max_norm),
                                    float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                    int(step), ptr(scratch), stream()), "dmt_adam_clip_step")


def mf_fwd(user, item, rating, Wu, Wi, bu, bi, bias, loss_kind, pu=None, pi=None, want_grad=True, colscale=None,
           add=None, want_q=False):
    lib = load()
    n = user.numel()
    dev = Wu.device
    pred = torch.empty(n, device=dev, dtype=torch.float32)
    dpred = torch.empty(n, device=dev, dtype=torch.float32) if want_grad else None
    q = torch.empty(n, Wu.shape[1], device=dev, dtype=torch.float32) if want_q else None
    sums = torch.empty(2, device=dev, dtype=torch.float32)
    scratch = torch.empty(lib.dmt_mf_scratch_floats(), device=dev, dtype=torch.float32)
    check(lib.dmt_mf_fwd(ptr(user), ptr(item), ptr(rating), n, ptr(Wu), ptr(Wi), ptr(bu), ptr(bi), ptr(bias), ptr(pu),
                         ptr(pi), ptr(colscale), ptr(add), Wu.shape[1], loss_kind, ptr(pred), ptr(dpred), ptr(q),
                         ptr(sums), ptr(scratch), stream()), "dmt_mf_fwd")
    if want_q:
        return pred, dpred, sums, q
    return pred, dpred, sums


def mf_bwd_table(other, W_other, b_other, p_side, dpred, scale, seg, n_rows, colscale=None):
    """Dense grads (dW [n_rows x H], db [n_rows]) of the table whose sorted segments are ``seg``."""
    perm, seg_key, seg_off, n_seg = seg
    H = W_other.shape[1]
    dW = torch.zeros(n_rows, H, device=W_other.device, dtype=torch.float32)
    db = torch.zeros(n_rows, device=W_other.device, dtype=torch.float32)
    check(load().dmt_mf_bwd_table(ptr(other), ptr(W_other), ptr(b_other), ptr(p_side), H, ptr(dpred), float(scale),
                                  ptr(perm), ptr(seg_key), ptr(seg_off), ptr(n_seg), min(n_rows, other.numel()),
                                  ptr(colscale), ptr(dW), ptr(db), stream()), "dmt_mf_bwd_table")
    return dW, db


def mf_bwd_side(idx, W, b, dpred, scale, colscale=None):
    H = W.shape[1]
    d_p = torch.empty(idx.numel(), H, device=W.device, dtype=torch.float32)
    check(load().dmt_mf_bwd_side(ptr(idx), idx.numel(), ptr(W), ptr(b), H, ptr(dpred), float(scale), ptr(colscale),
                                 ptr(d_p), stream()), "dmt_mf_bwd_side")
    return d_p


def embed_fwd(idx, W, b, out, col_off):
    check(load().dmt_embed_fwd(ptr(idx), idx.numel(), ptr(W), ptr(b), W.shape[1], ptr(out), out.shape[1], col_off,
                               stream()), "dmt_embed_fwd")


def embed_bwd(dOut, col_off, H, seg, n_rows):
    perm, seg_key, seg_off, n_seg = seg
    dW = torch.zeros(n_rows, H, device=dOut.device, dtype=torch.float32)
    db = torch.zeros(n_rows, device=dOut.device, dtype=torch.float32)
    check(load().dmt_embed_bwd(ptr(dOut), dOut.shape[1], col_off, H, ptr(perm), ptr(seg_key), ptr(seg_off), ptr(n_seg),
                               min(n_rows, dOut.shape[0]), ptr(dW), ptr(db), stream()), "dmt_embed_bwd")
    return dW, db


def weighted_colsum(g, Q, scale=1.0):
    lib = load()
    n, width = Q.shape
    out = torch.empty(width, device=Q.device, dtype=torch.float32)
    scratch = torch.empty(lib.dmt_weighted_colsum_scratch_floats(width), device=Q.device, dtype=torch.float32)
    check(lib.dmt_weighted_colsum(ptr(g), float(scale), ptr(Q), n, width, Q.shape[1], ptr(out), ptr(scratch),
                                  stream()), "dmt_weighted_colsum")
    return out


def loss_fwd(pred, y, loss_kind, want_grad=True):
    lib = load()
    n = pred.numel()
    dpred = torch.empty_like(pred) if want_grad else None
    sums = torch.empty(2, device=pred.device, dtype=torch.float32)
    scratch = torch.empty(lib.dmt_mf_scratch_floats(), device=pred.device, dtype=torch.float32)
    check(lib.dmt_loss_fwd(ptr(pred), ptr(y), n, loss_kind, ptr(dpred), ptr(sums), ptr(scratch), stream()),
          "dmt_loss_fwd")
    return dpred, sums


def dense_fwd(X, W, b, act, keep=None, keep_scale=1.0):
    m, k = X.shape
    n = W.shape[0]
    Y = torch.empty(m, n, device=X.device, dtype=torch.float32)
    Y_pre = torch.empty_like(Y) if keep is not None else None
    check(load().dmt_dense_fwd(ptr(X), ptr(W), ptr(b), ptr(Y), ptr(Y_pre), ptr(keep), float(keep_scale), m, n, k, act,
                               stream()), "dmt_dense_fwd")
    return Y, Y_pre


def dense_bwd_x(dY, W, A_prev, act_prev, keep=None, keep_scale=1.0):
    m, n = dY.shape
    k = W.shape[1]
    dX = torch.empty(m, k, device=dY.device, dtype=torch.float32)
    check(load().dmt_dense_bwd_x(ptr(dY), ptr(W), ptr(A_prev), ptr(keep), float(keep_scale), ptr(dX), m, n, k,
                                 act_prev, stream()), "dmt_dense_bwd_x")
    return dX


def dense_bwd_w(dY, X, want_bias=True):
    m, n = dY.shape
    k = X.shape[1]
    dW = torch.empty(n, k, device=dY.device, dtype=torch.float32)
    db = torch.empty(n, device=dY.device, dtype=torch.float32) if want_bias else None
    check(load().dmt_dense_bwd_w(ptr(dY), ptr(X), ptr(dW), ptr(db), m, n, k, stream()), "dmt_dense_bwd_w")
    return dW, db


def dense_fwd_tc(X, W, b, act, keep=None, keep_scale=1.0, passes=3):
    """tcgen05 form of :func:`dense_fwd` (3xTF32 when passes == 3)."""
    m, k = X.shape
    n = W.shape[0]
    Y = torch.empty(m, n, device=X.device, dtype=torch.float32)
    Y_pre = torch.empty_like(Y) if keep is not None else None
    check(load().dmt_dense_fwd_tc(ptr(X), ptr(W), ptr(b), ptr(Y), ptr(Y_pre), ptr(keep), float(keep_scale), m, n, k,
                                  act, passes, stream()), "dmt_dense_fwd_tc")
    return Y, Y_pre


def dense_bwd_x_tc(dY, W, A_prev, act_prev, keep=None, keep_scale=1.0, passes=3):
    m, n = dY.shape
    k = W.shape[1]
    dX = torch.empty(m, k, device=dY.device, dtype=torch.float32)
    check(load().dmt_dense_bwd_x_tc(ptr(dY), ptr(W), ptr(A_prev), ptr(keep), float(keep_scale), ptr(dX), m, n, k,
                                    act_prev, passes, stream()), "dmt_dense_bwd_x_tc")
    return dX


def dense_bwd_w_tc(dY, X, want_bias=True, passes=3):
    m, n = dY.shape
    k = X.shape[1]
    dW = torch.empty(n, k, device=dY.device, dtype=torch.float32)
    db = torch.empty(n, device=dY.device, dtype=torch.float32) if want_bias else None
    check(load().dmt_dense_bwd_w_tc(ptr(dY), ptr(X), ptr(dW), ptr(db), m, n, k, passes, stream()),
          "dmt_dense_bwd_w_tc")
    return dW, db


def ae_decoder_tc(rows, indptr, indices, target, A3, W4, b4, loss_kind, nnz, train, tanh_deriv=True, passes=3):
    """Tensor-core decoder (dmt_ae_decoder_tc). Returns pred, gout, dZ3, dW4, db4, loss_sum (train) aligned like
    :func:`ae_decoder_fwd`; the CSR's column indices must ascend inside every row."""
    lib = load()
    dev = A3.device
    H = A3.shape[1]
    n_dec = W4.shape[0]
    n_rows = rows.numel()
    pred = torch.zeros(nnz, device=dev, dtype=torch.float32)
    gout = dz3 = dW4 = db4 = loss_rows = n_t = scratch = None
    if train:
        gout = torch.zeros(nnz, device=dev, dtype=torch.float32)
        dz3 = torch.empty_like(A3)
        dW4 = torch.empty_like(W4)
        db4 = torch.empty(n_dec, device=dev, dtype=torch.float32)
        loss_rows = torch.empty(max(n_rows, 1), device=dev, dtype=torch.float32)
        ip = indptr.to(torch.int64)
        r = rows.to(torch.int64)
        n_t = (ip[r + 1] - ip[r]).sum().to(torch.int32).reshape(1)
        scratch = torch.empty(max(1, lib.dmt_ae_decoder_tc_scratch_floats(n_rows, n_dec, H)), device=dev,
                              dtype=torch.float32)
    check(lib.dmt_ae_decoder_tc(ptr(rows), n_rows, ptr(indptr), ptr(indices), ptr(target), ptr(A3), ptr(W4), ptr(b4), H,
                                n_dec, loss_kind, ptr(n_t), passes, ptr(pred), ptr(gout), ptr(dz3), ptr(dW4), ptr(db4),
                                ptr(loss_rows), int(bool(tanh_deriv)), ptr(scratch), stream()), "dmt_ae_decoder_tc")
    return pred, gout, dz3, dW4, db4, loss_rows, n_t


def ae_encoder_fwd(rows, indptr, indices, val, W1t, b1):
    H = W1t.shape[1]
    A1 = torch.empty(rows.numel(), H, device=W1t.device, dtype=torch.float32)
    check(load().dmt_ae_encoder_fwd(ptr(rows), rows.numel(), ptr(indptr), ptr(indices), ptr(val), ptr(W1t), ptr(b1),
                                    H, ptr(A1), stream()), "dmt_ae_encoder_fwd")
    return A1


def ae_decoder_fwd(rows, indptr, indices, target, A3, W4, b4, loss_kind, nnz, train, tanh_deriv=True):
    """pred is aligned with the target CSR (length nnz). Train mode also returns gout (same alignment),
    dZ3 [rows x H] and the per-row loss sums."""
    dev = A3.device
    H = A3.shape[1]
    pred = torch.zeros(nnz, device=dev, dtype=torch.float32)
    gout = dz3 = loss_rows = n_t = None
    if train:
        gout = torch.zeros(nnz, device=dev, dtype=torch.float32)
        dz3 = torch.empty_like(A3)
        loss_rows = torch.empty(rows.numel(), device=dev, dtype=torch.float32)
        ip = indptr.to(torch.int64)
        r = rows.to(torch.int64)
        n_t = (ip[r + 1] - ip[r]).sum().to(torch.int32).reshape(1)
    check(load().dmt_ae_decoder_fwd(ptr(rows), rows.numel(), ptr(indptr), ptr(indices), ptr(target), ptr(A3), ptr(W4),
                                    ptr(b4), H, loss_kind, ptr(n_t), ptr(pred), ptr(gout), ptr(dz3), ptr(loss_rows),
                                    int(bool(tanh_deriv)), stream()), "dmt_ae_decoder_fwd")
    return pred, gout, dz3, loss_rows, n_t


def eval_blocks(indptr, pred, target, n_rows, block_rows, loss_kind, block_k=None):
    """Per-block sums [n_blocks x 3] = (loss, squared error, sum of per-row NDCG) — dmt_eval_blocks."""
    n_blocks = (n_rows + block_rows - 1) // block_rows
    out = torch.zeros(n_blocks, 3, device=pred.device, dtype=torch.float32)
    check(load().dmt_eval_blocks(ptr(indptr), ptr(pred), ptr(target), n_rows, block_rows, loss_kind,
                                 int(block_k is not None), ptr(block_k), ptr(out), stream()), "dmt_eval_blocks")
    return out


def privacy(y, mode, param, seed, out=None):
    """make_privacy on the device (dmt_privacy). Returns (perturbed vector, [a, b] quantile pair)."""
    lib = load()
    n = y.numel()
    out = torch.empty_like(y) if out is None else out
    q = torch.empty(2, device=y.device, dtype=torch.float32)
    nbytes = lib.dmt_privacy_temp_bytes(n)
    temp = torch.empty(nbytes, device=y.device, dtype=torch.uint8)
    check(lib.dmt_privacy(ptr(y), n, {"dp": 0, "ip": 1}[mode], float(param), int(seed) & (2 ** 64 - 1), ptr(out), ptr(q),
                          ptr(temp), nbytes, stream()), "dmt_privacy")
    return out, q


class Group:
    """Handle of an organization group (dmt_group_*): lockstep training of several organizations, one launch per
    step kernel for all of them."""

    def __init__(self, orgs):
        lib = load()
        self._lib = lib
        self.orgs = list(orgs)
        arr = (C.c_void_p * len(self.orgs))(*[o.h for o in self.orgs])
        self.h = C.c_void_p()
        check(lib.dmt_group_create(C.byref(self.h), arr, len(self.orgs), None), "dmt_group_create")

    def train(self, rows, row_off, n_rows_total, n_batches, n_t, n_d, seeds, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
              weight_decay=5e-4, max_norm=1.0, batch_loss=None):
        n = len(self.orgs)
        P = C.c_void_p
        rows_a = (P * n)(*[ptr(r) for r in rows])
        off_a = (P * n)(*[ptr(r) for r in row_off])
        nt_a = (C.c_int64 * n)(*[int(x) for x in n_t])
        nd_a = (C.c_int64 * n)(*[int(x) for x in n_d])
        seed_a = (C.c_uint64 * n)(*[int(x) for x in seeds])
        loss_a = (P * n)(*[ptr(t) for t in batch_loss]) if batch_loss is not None else None
        check(self._lib.dmt_group_wait_stream(self.h, stream()), "dmt_group_wait_stream")
        check(self._lib.dmt_group_train(self.h, rows_a, off_a, int(n_rows_total), int(n_batches), nt_a, nd_a, seed_a,
                                        float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                        float(max_norm), loss_a), "dmt_group_train")

    def sync(self):
        check(self._lib.dmt_group_sync(self.h), "dmt_group_sync")

    def close(self):
        if self.h:
            self._lib.dmt_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Org:
    """Handle of the device-resident organization engine (dmt_org_*)."""

    def __init__(self, n_rows, n_enc, n_dec, H1, H2, d_csr, t_csr, batch_rows, loss_kind, own_stream=True,
                 plan_epochs=1):
        lib = load()
        self._lib = lib
        self.d_csr = d_csr  # (indptr, indices, val) int32/int32/float32 CUDA tensors, kept alive here
        self.t_csr = t_csr  # (indptr, indices)
        self.h = C.c_void_p()
        st = None if own_stream else stream()
        check(lib.dmt_org_create(C.byref(self.h), n_rows, n_enc, n_dec, H1, H2, ptr(d_csr[0]), ptr(d_csr[1]),
                                 ptr(d_csr[2]), d_csr[1].numel(), ptr(t_csr[0]), ptr(t_csr[1]), t_csr[1].numel(),
                                 batch_rows, loss_kind, int(plan_epochs), st), "dmt_org_create")
        self.n_params = lib.dmt_org_num_params(self.h)
        self.H2 = H2
        self._target = None

    def set_decoder_mode(self, mode, passes=3):
        """mode 'gather' (SDDMM + segmented reductions) or 'tc' (tcgen05 GEMMs; sorted column indices needed)."""
        code = {"gather": 0, "tc": 1}[mode]
        check(self._lib.dmt_org_set_decoder_mode(self.h, code, int(passes)), "dmt_org_set_decoder_mode")
        self.decoder_mode = mode

    def set_decoder_blocks(self, blocks):
        """Grid of the decoder chunk kernel (0: two blocks per SM); fewer blocks pay with many organizations per GPU."""
        check(self._lib.dmt_org_set_decoder_blocks(self.h, int(blocks)), "dmt_org_set_decoder_blocks")

    def set_row_tile(self, rows):
        """Batch rows per CTA of the fused step's row kernels (dmt_org_set_row_tile): 4 for ranks with few organizations."""
        check(self._lib.dmt_org_set_row_tile(self.h, int(rows)), "dmt_org_set_row_tile")

    def set_pdl(self, on):
        """Programmatic dependent launch between the kernels of the fused step (dmt_org_set_pdl)."""
        check(self._lib.dmt_org_set_pdl(self.h, int(bool(on))), "dmt_org_set_pdl")

    def set_gather_mode(self, mode):
        """'ldg' (register loads, default) or 'bulk' (one cp.async.bulk copy per row into shared-memory rings,
        csrc/bulk.cuh; measured slower, kept as the parity-tested alternative)."""
        check(self._lib.dmt_org_set_gather_mode(self.h, {"ldg": 0, "bulk": 1}[mode]), "dmt_org_set_gather_mode")

    def gather_mode(self):
        return "bulk" if self._lib.dmt_org_gather_mode(self.h) == 1 else "ldg"

    def set_step_mode(self, mode):
        """'fused' (six launches per batch, csrc/fused.cu) or 'classic' (one kernel per layer / reduction)."""
        check(self._lib.dmt_org_set_step_mode(self.h, {"classic": 0, "fused": 1}[mode]), "dmt_org_set_step_mode")

    def step_mode(self):
        return "fused" if self._lib.dmt_org_step_mode(self.h) else "classic"

    def set_fanout(self, on):
        """Backward pass of a step as parallel graph branches (pays with few organizations per GPU)."""
        check(self._lib.dmt_org_set_fanout(self.h, int(bool(on))), "dmt_org_set_fanout")

    def set_params(self, flat):
        check(self._lib.dmt_org_set_params(self.h, ptr(flat)), "dmt_org_set_params")

    def get_params(self, out=None):
        out = torch.empty(self.n_params, device=self.d_csr[0].device, dtype=torch.float32) if out is None else out
        check(self._lib.dmt_org_get_params(self.h, ptr(out)), "dmt_org_get_params")
        return out

    def set_target(self, t_val):
        self._target = t_val
        check(self._lib.dmt_org_set_target(self.h, ptr(t_val)), "dmt_org_set_target")

    def train_epoch(self, rows, row_off, n_t_entries, n_d_entries, keep=None, seed=0, lr=1e-3, betas=(0.9, 0.999),
                    eps=1e-8, weight_decay=5e-4, max_norm=1.0, epoch_loss=None):
        nb = row_off.numel() - 1
        # the row lists, masks and loss buffer were usually produced on torch's current stream just before this call:
        # order the organization's stream behind it (a fill kernel of `epoch_loss` landing after the epoch would erase it)
        self.wait_current()
        check(self._lib.dmt_org_train_epoch(self.h, ptr(rows), ptr(row_off), rows.numel(), nb, int(n_t_entries),
                                            int(n_d_entries), ptr(keep), int(seed), float(lr), float(betas[0]),
                                            float(betas[1]), float(eps), float(weight_decay), float(max_norm),
                                            ptr(epoch_loss)), "dmt_org_train_epoch")

    def predict(self, d_csr, t_csr, n_rows, out):
        self.wait_current()  # `out` and the CSR arrays may still be in flight on torch's current stream
        check(self._lib.dmt_org_predict(self.h, ptr(d_csr[0]), ptr(d_csr[1]), ptr(d_csr[2]), ptr(t_csr[0]),
                                        ptr(t_csr[1]), n_rows, ptr(out)), "dmt_org_predict")
        return out

    def sync(self):
        check(self._lib.dmt_org_sync(self.h), "dmt_org_sync")

    def wait_current(self):
        """The organization's stream waits for work already enqueued on torch's current stream."""
        check(self._lib.dmt_org_wait_stream(self.h, stream()), "dmt_org_wait_stream")

    def signal_current(self):
        """torch's current stream waits for work already enqueued on the organization's stream."""
        check(self._lib.dmt_org_signal_stream(self.h, stream()), "dmt_org_signal_stream")

    PROFILE_CLASSES = ["encoder_spmm", "dense_fwd", "grad_zero", "decoder_loss_dz3", "dw4_segments", "dense_bwd",
                       "dw1_segments", "grad_norm", "clip_adam"]

    def profile_step(self, b=0, reps=20):
        """{kernel class: average ms} for one training step on batch b of the current plan (dmt_org_profile_step)."""
        n = self._lib.dmt_org_profile_classes()
        buf = (C.c_float * n)()
        check(self._lib.dmt_org_profile_step(self.h, int(b), int(reps), buf), "dmt_org_profile_step")
        vals = [float(x) for x in buf]
        if self.step_mode() == "fused":  # six launches (+ the dW4 branch): csrc/fused.cu
            names = {"encoder_spmm": "fwd_rows", "decoder_loss_dz3": "decoder_loss_dz3", "dw4_segments": "dw4_segments",
                     "dw1_segments": "bwd_rows", "dense_bwd": "grad_phase", "grad_norm": "grad_norm",
                     "clip_adam": "clip_adam"}
            return {names[k]: v for k, v in zip(self.PROFILE_CLASSES, vals) if k in names}
        return dict(zip(self.PROFILE_CLASSES, vals))

    def cuda_stream(self):
        return self._lib.dmt_org_stream(self.h)

    def close(self):
        if self.h:
            self._lib.dmt_org_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
