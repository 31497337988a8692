"""torch.autograd bridges onto the C-ABI kernels (module-level path of the drop-in models).

The reference's drivers own ``loss.backward()``, ``clip_grad_norm_`` and the optimizer for MF/MLP/NMF
(src/train_recsys_joint.py:118-134), so the models must stay ``nn.Module``s whose parameters receive dense
``.grad``s. Each Function below runs the forward AND the backward arithmetic in libdmt_b200 kernels; torch
only carries tensors between them. CUDA tensors only: there is no CPU path.
"""
import torch

from dmtcdr_b200 import native


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise native.NativeError("dmtcdr_b200 models run on CUDA tensors only (got a CPU tensor); "
                                     "move the model and the batch to cfg['device']='cuda'")


_SORT_CACHE = {}


def sorted_segments(idx, bound):
    """native.sort_segments memoised on the index tensor: the NCF head and the tower input of one step gather with the
    same user / item vectors, so each is sorted once per step instead of once per autograd node (the cache holds the
    tensors of the latest step only)."""
    key = (idx.data_ptr(), idx.numel(), int(bound), idx._version)
    hit = _SORT_CACHE.get(key)
    if hit is not None and hit[0] is idx:
        return hit[1]
    if len(_SORT_CACHE) >= 8:
        _SORT_CACHE.clear()
    seg = native.sort_segments(idx, bound)
    _SORT_CACHE[key] = (idx, seg)
    return seg


def _scaled(t, f, inplace=True):
    """t * f with f = dloss / n a device scalar: no host read of dloss (float(dloss) would drain the stream every
    step). Fresh gradient buffers are scaled in place."""
    if t is None:
        return None
    return t.mul_(f) if inplace else t * f


def batch_csr(row_ids, cols, vals, rows_sorted):
    """Batch-local CSR of COO triples: local row = rank of row_ids in rows_sorted; stable within a row."""
    r = torch.searchsorted(rows_sorted, row_ids)
    order = torch.argsort(r, stable=True)
    counts = torch.bincount(r, minlength=rows_sorted.numel())
    indptr = torch.zeros(rows_sorted.numel() + 1, dtype=torch.int32, device=row_ids.device)
    indptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return (indptr, cols[order].to(torch.int32).contiguous(), vals[order].contiguous() if vals is not None else None,
            order, r[order].to(torch.int32).contiguous())


class SparseEncoderFn(torch.autograd.Function):
    """A1 = tanh(X W^T + b) with X given as batch-local CSR (reference src/models/ae.py:101-110)."""

    @staticmethod
    def forward(ctx, weight, bias, indptr, indices, val, ent_row):
        _need_cuda(weight, indptr)
        W1t = weight.t().contiguous()
        n_rows = indptr.numel() - 1
        rows = torch.arange(n_rows, dtype=torch.int32, device=weight.device)
        A1 = native.ae_encoder_fwd(rows, indptr, indices, val, W1t, bias.contiguous())
        ctx.save_for_backward(A1, indices, val, ent_row)
        ctx.n_cols = weight.shape[1]
        return A1

    @staticmethod
    def backward(ctx, dA1):
        A1, indices, val, ent_row = ctx.saved_tensors
        dZ1 = (dA1 * (1.0 - A1 * A1)).contiguous()
        H = A1.shape[1]
        dW1t = torch.zeros(ctx.n_cols, H, device=A1.device, dtype=torch.float32)
        if indices.numel() > 0:
            seg = native.sort_segments(indices, ctx.n_cols)
            native.segment_reduce_rows(*seg, min(ctx.n_cols, indices.numel()), val, ent_row, dZ1, dW1t)
        return dW1t.t(), dZ1.sum(0), None, None, None, None


class DenseFn(torch.autograd.Function):
    """Y = act(X W^T + b) [* keep * scale]; act 0 none / 1 tanh / 2 relu (reference nn.Linear + Tanh/ReLU blocks)."""

    @staticmethod
    def forward(ctx, X, weight, bias, act, keep, scale):
        _need_cuda(X, weight)
        X = X.contiguous()
        Y, Y_pre = native.dense_fwd(X, weight.contiguous(), bias.contiguous() if bias is not None else None, act,
                                    keep, scale)
        ctx.save_for_backward(X, weight, Y_pre if keep is not None else Y, keep)
        ctx.act, ctx.scale, ctx.has_bias = act, scale, bias is not None
        return Y

    @staticmethod
    def backward(ctx, dY):
        X, weight, pre, keep = ctx.saved_tensors
        dZ = dY
        if keep is not None:
            dZ = dZ * (keep.to(dY.dtype) * ctx.scale)
        if ctx.act == 1:
            dZ = dZ * (1.0 - pre * pre)
        elif ctx.act == 2:
            dZ = dZ * (pre > 0).to(dY.dtype)
        dZ = dZ.contiguous()
        dX = native.dense_bwd_x(dZ, weight.contiguous(), None, 0) if ctx.needs_input_grad[0] else None
        dW, db = native.dense_bwd_w(dZ, X, want_bias=ctx.has_bias)
        return dX, dW, db, None, None, None


class SparseDecoderLossFn(torch.autograd.Function):
    """pred_e = A3[r_e] . W4[c_e] + b4[c_e] at the batch's target entries, fused with the mean loss
    (reference src/models/ae.py:135-142,153-156). Returns (pred in CSR order, loss)."""

    @staticmethod
    def forward(ctx, A3, weight, bias, indptr, indices, target, ent_row, loss_kind):
        _need_cuda(A3, weight)
        n_rows = indptr.numel() - 1
        rows = torch.arange(n_rows, dtype=torch.int32, device=A3.device)
        train = any(ctx.needs_input_grad[:3])  # grad mode is off inside forward(); this is the reliable signal
        ctx.set_materialize_grads(False)
        A3c, Wc, bc = A3.contiguous(), weight.contiguous(), bias.contiguous()
        nnz = indices.numel()
        if nnz == 0:
            raise native.NativeError("AE batch without target entries")
        pred, gout, dA3, loss_rows, n_t = native.ae_decoder_fwd(rows, indptr, indices, target, A3c, Wc, bc, loss_kind,
                                                                nnz, True, tanh_deriv=False)
        loss = loss_rows.sum() / nnz
        if train:
            ctx.save_for_backward(A3c, gout, dA3, indices, ent_row)
            ctx.n_cols = weight.shape[0]
        ctx.train = train
        ctx.mark_non_differentiable(pred)
        return pred, loss

    @staticmethod
    def backward(ctx, dpred, dloss):
        if dloss is None:
            return (None,) * 8
        A3, gout, dA3, indices, ent_row = ctx.saved_tensors
        H = A3.shape[1]
        dW4 = torch.zeros(ctx.n_cols, H, device=A3.device, dtype=torch.float32)
        db4 = torch.zeros(ctx.n_cols, device=A3.device, dtype=torch.float32)
        seg = native.sort_segments(indices, ctx.n_cols)
        native.segment_reduce_rows(*seg, min(ctx.n_cols, indices.numel()), gout, ent_row, A3, dW4, db4)
        return dA3 * dloss, dW4 * dloss, db4 * dloss, None, None, None, None, None


class MFFn(torch.autograd.Function):
    """MF forward + loss, dense embedding gradients by sort + segmented reduction (reference src/models/mf.py:57-93)."""

    @staticmethod
    def forward(ctx, user, item, rating, Wu, Wi, bu, bi, bias, pu, pi, loss_kind):
        _need_cuda(user, Wu)
        Wu_, Wi_ = Wu.contiguous(), Wi.contiguous()
        bu_, bi_ = bu.reshape(-1).contiguous(), bi.reshape(-1).contiguous()
        pu_ = pu.contiguous() if pu is not None else None
        pi_ = pi.contiguous() if pi is not None else None
        train = any(ctx.needs_input_grad)
        pred, dpred, sums = native.mf_fwd(user, item, rating, Wu_, Wi_, bu_, bi_, bias.contiguous(), loss_kind, pu_,
                                          pi_, want_grad=train)
        n = user.numel()
        loss = sums[0] / n
        ctx.set_materialize_grads(False)
        if train:
            ctx.save_for_backward(user, item, Wu_, Wi_, bu_, bi_, pu_, pi_, dpred, sums)
        ctx.mark_non_differentiable(pred)
        return pred, loss

    @staticmethod
    def backward(ctx, dpred_in, dloss):
        if dloss is None:
            return (None,) * 11
        user, item, Wu, Wi, bu, bi, pu, pi, dpred, sums = ctx.saved_tensors
        n = user.numel()
        seg_u = sorted_segments(user, Wu.shape[0])
        seg_i = sorted_segments(item, Wi.shape[0])
        dWu, dbu = native.mf_bwd_table(item, Wi, bi, pu, dpred, 1.0, seg_u, Wu.shape[0])
        dWi, dbi = native.mf_bwd_table(user, Wu, bu, pi, dpred, 1.0, seg_i, Wi.shape[0])
        dpu = native.mf_bwd_side(user, Wu, bu, dpred, 1.0) if pu is not None else None
        dpi = native.mf_bwd_side(item, Wi, bi, dpred, 1.0) if pi is not None else None
        f = dloss / n
        sc = lambda t: _scaled(t, f)  # noqa: E731
        return (None, None, None, sc(dWu), sc(dWi), sc(dbu).view(-1, 1), sc(dbi).view(-1, 1),
                _scaled(sums[1], f, False).reshape(1), sc(dpu), sc(dpi), None)


class EmbedCatFn(torch.autograd.Function):
    """[W_u[user] + b_u[user], W_i[item] + b_i[item]] written straight into one [n x 2H] tower-input buffer
    (reference src/models/mlp.py:96, nmf.py:127); backward = sort-by-index + segmented row sums."""

    @staticmethod
    def forward(ctx, user, item, Wu, bu, Wi, bi):
        _need_cuda(user, Wu)
        H = Wu.shape[1]
        out = torch.empty(user.numel(), 2 * H, device=Wu.device, dtype=torch.float32)
        native.embed_fwd(user, Wu.contiguous(), bu.reshape(-1).contiguous(), out, 0)
        native.embed_fwd(item, Wi.contiguous(), bi.reshape(-1).contiguous(), out, H)
        ctx.save_for_backward(user, item)
        ctx.sizes = (Wu.shape[0], Wi.shape[0], H)
        return out

    @staticmethod
    def backward(ctx, dOut):
        user, item = ctx.saved_tensors
        nu, ni, H = ctx.sizes
        dOut = dOut.contiguous()
        dWu, dbu = native.embed_bwd(dOut, 0, H, sorted_segments(user, nu), nu)
        dWi, dbi = native.embed_bwd(dOut, H, H, sorted_segments(item, ni), ni)
        return None, None, dWu, dbu.view(-1, 1), dWi, dbi.view(-1, 1)


class GMFLossFn(torch.autograd.Function):
    """NCF head: pred = sum_d a_d * (u~ (i~ [+pu]) [+ i~ pi])_d + add, fused with the loss
    (reference src/models/nmf.py:126-146 with the affine layer split into its tower part `add` and GMF part `a`)."""

    @staticmethod
    def forward(ctx, user, item, rating, Wu, Wi, bu, bi, pu, pi, colscale, add, loss_kind):
        _need_cuda(user, Wu)
        Wu_, Wi_ = Wu.contiguous(), Wi.contiguous()
        bu_, bi_ = bu.reshape(-1).contiguous(), bi.reshape(-1).contiguous()
        pu_ = pu.contiguous() if pu is not None else None
        pi_ = pi.contiguous() if pi is not None else None
        cs = colscale.contiguous()
        train = any(ctx.needs_input_grad)
        res = native.mf_fwd(user, item, rating, Wu_, Wi_, bu_, bi_, None, loss_kind, pu_, pi_, want_grad=train,
                            colscale=cs, add=add.contiguous(), want_q=train)
        pred, dpred, sums = res[0], res[1], res[2]
        loss = sums[0] / user.numel()
        ctx.set_materialize_grads(False)
        if train:
            ctx.save_for_backward(user, item, Wu_, Wi_, bu_, bi_, pu_, pi_, cs, dpred, res[3])
        ctx.mark_non_differentiable(pred)
        return pred, loss

    @staticmethod
    def backward(ctx, dpred_in, dloss):
        if dloss is None:
            return (None,) * 12
        user, item, Wu, Wi, bu, bi, pu, pi, cs, dpred, q = ctx.saved_tensors
        n = user.numel()
        seg_u = sorted_segments(user, Wu.shape[0])
        seg_i = sorted_segments(item, Wi.shape[0])
        dWu, dbu = native.mf_bwd_table(item, Wi, bi, pu, dpred, 1.0, seg_u, Wu.shape[0], cs)
        dWi, dbi = native.mf_bwd_table(user, Wu, bu, pi, dpred, 1.0, seg_i, Wi.shape[0], cs)
        dpu = native.mf_bwd_side(user, Wu, bu, dpred, 1.0, cs) if pu is not None else None
        dpi = native.mf_bwd_side(item, Wi, bi, dpred, 1.0, cs) if pi is not None else None
        dcs = native.weighted_colsum(dpred, q, 1.0)
        f = dloss / n
        sc = lambda t: _scaled(t, f)  # noqa: E731
        return (None, None, None, sc(dWu), sc(dWi), sc(dbu).view(-1, 1), sc(dbi).view(-1, 1), sc(dpu), sc(dpi), sc(dcs),
                _scaled(dpred, f, False), None)


class LossFn(torch.autograd.Function):
    """models.loss_fn (mean) with its gradient, one kernel (reference src/models/utils.py:7-14)."""

    @staticmethod
    def forward(ctx, pred, target, loss_kind):
        _need_cuda(pred)
        dpred, sums = native.loss_fwd(pred.contiguous(), target.contiguous(), loss_kind, ctx.needs_input_grad[0])
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(dpred)
        ctx.n = pred.numel()
        return sums[0] / ctx.n

    @staticmethod
    def backward(ctx, dloss):
        dpred, = ctx.saved_tensors
        return dpred * (dloss / ctx.n), None, None


def dense(X, weight, bias, act=0, keep=None, scale=1.0):
    return DenseFn.apply(X, weight, bias, act, keep, scale)
