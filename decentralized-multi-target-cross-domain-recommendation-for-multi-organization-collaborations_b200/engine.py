"""Device-resident MTAL engine: the data layout in HBM and the host-side sequencing of one assistance round.

Layout (per rank):
  * one canonical global CSR per split (rows = aligned entity, columns = ALL items/users): ``indptr``/``indices``
    int32 + the ground truth ``y`` fp32; every per-entry vector of the round (F_t, residual, each organization's
    prediction) is a plain fp32 array aligned with that CSR's storage order — the positional contract the
    reference relies on (src/assist.py:45-46,94-117).
  * ``O[split]``: [K x nnz] organization-major matrix of the K organizations' predictions (the payload of the
    per-round exchange; rows of other ranks arrive by allgather).
  * per organization: its data column block as CSR (local column ids) + an ``native.Org`` handle that owns the
    parameters, Adam moments, activations, epoch plan and the captured per-epoch CUDA graph.
"""
from __future__ import annotations

import math
import os
import zlib

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import native


XFER = {"h2d": 0, "d2h": 0}  # bytes moved between host and device through this package (bench.py reads it)


def to_dev(x, device):
    t = torch.from_numpy(x) if isinstance(x, np.ndarray) else x
    XFER["h2d"] += t.numel() * t.element_size()
    return t.to(device)


_STAGING = {}  # device index -> persistent pinned staging buffer (uint8)


def to_host(t):
    """Device -> host through ONE persistent pinned staging buffer per device, then a copy into a fresh pageable
    tensor; counted in XFER. (Allocating pinned memory per call — what a fresh pin_memory tensor does whenever torch's
    host cache has no free block, e.g. while the previous round's outputs are still referenced — costs cudaHostAlloc
    calls that were measured to stall a round by 0.2-0.9 s now and then; the staging copy costs ~0.4 ms per 4 MB.)"""
    XFER["d2h"] += t.numel() * t.element_size()
    if not t.is_cuda or t.numel() < (1 << 14):
        return t.cpu()
    nbytes = t.numel() * t.element_size()
    key = t.device.index
    st = _STAGING.get(key)
    if st is None or st.numel() < nbytes:
        st = torch.empty(max(nbytes, 16 << 20), dtype=torch.uint8, pin_memory=True)
        _STAGING[key] = st
    view = st[:nbytes].view(t.dtype).view(t.shape)
    view.copy_(t.contiguous(), non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    out = torch.empty(t.shape, dtype=t.dtype)
    out.copy_(view)
    return out


class DeviceCSR:
    """CSR on the device (int32 indices) with a host copy of indptr for sizing decisions."""

    def __init__(self, m, device="cuda", with_values=True):
        m = m.tocsr()
        if m.nnz >= 2 ** 31 - 2:
            raise ValueError("nnz must fit int32")
        self.shape = m.shape
        self.nnz = int(m.nnz)
        # ascending column indices inside every row: what the tensor-core decoder's CSR windows need (the storage
        # order itself is never changed: it is the positional contract of src/assist.py:45-46)
        self.sorted = bool(m.has_sorted_indices)
        self.indptr_host = np.asarray(m.indptr, dtype=np.int64)
        self.indices_host = np.asarray(m.indices)
        self.indptr = to_dev(self.indptr_host.astype(np.int32), device)
        self.indices = to_dev(self.indices_host.astype(np.int32), device)
        self.data = to_dev(np.asarray(m.data, dtype=np.float32), device) if with_values else None

    @property
    def row_len(self):
        return np.diff(self.indptr_host)

    def triple(self):
        return (self.indptr, self.indices, self.data)

    def pair(self):
        return (self.indptr, self.indices)


def csr_key(m):
    """Content key of a scipy CSR's structure (used to reuse device uploads across API calls)."""
    return (m.shape, int(m.nnz), zlib.crc32(np.ascontiguousarray(m.indptr).view(np.uint8)),
            zlib.crc32(np.ascontiguousarray(m.indices).view(np.uint8)))


def index_batches(n, batch_size, shuffle):
    """Row-id batches of one pass of the reference's DataLoader (src/data.py:68-81). A real DataLoader over
    ``range(n)`` is iterated so that torch's global generator is consumed exactly as in the reference: one base-seed
    draw per iterator plus the RandomSampler's seed draw when shuffling."""
    return [np.asarray(b, dtype=np.int64) for b in DataLoader(range(n), batch_size=batch_size, shuffle=shuffle,
                                                              collate_fn=lambda x: x)]


def fast_perm_batches(n, batch_size, generator=None):
    """Shuffled row-id batches without the DataLoader machinery (production mode)."""
    perm = torch.randperm(n, generator=generator).numpy()
    return [perm[s:s + batch_size] for s in range(0, n, batch_size)]


class EpochLayout:
    """Host-side description of one local epoch for ``dmt_org_train_epoch``: per batch the sorted row ids with
    entries (the reference's ``total_user``, src/models/ae.py:101), offsets, entry totals, and which batches carry
    data (batches without data are skipped, src/organization.py:153-155)."""

    def __init__(self, batches, d_len, t_len):
        rows, off, active, b_rows = [], [0], [], []
        for b in batches:
            b = np.sort(np.asarray(b, dtype=np.int64))
            b = b[(d_len[b] + t_len[b]) > 0]
            rows.append(b)
            off.append(off[-1] + len(b))
            active.append(bool(d_len[b].sum() > 0))
            b_rows.append(len(b))
        self.rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        self.row_off = np.asarray(off, dtype=np.int32)
        self.active = active
        self.batch_rows = b_rows
        self.n_t = int(t_len[self.rows].sum())
        self.n_d = int(d_len[self.rows].sum())
        self.d_per_batch = [int(d_len[r].sum()) for r in rows]


def flat_from_state_dict(sd, device):
    """Reference AE state_dict -> the engine's flat layout W1t b1 W2 b2 W3 b3 W4 b4 (include/dmt_b200.h)."""
    g = lambda k: to_dev(sd[k].detach().to(torch.float32), device)
    return torch.cat([g("encoder_linear.weight").t().contiguous().view(-1), g("encoder_linear.bias"),
                      g("encoder.blocks.0.weight").reshape(-1), g("encoder.blocks.0.bias"),
                      g("decoder.blocks.0.weight").reshape(-1), g("decoder.blocks.0.bias"),
                      g("decoder_linear.weight").reshape(-1), g("decoder_linear.bias")]).contiguous()


def state_dict_from_flat(flat, n_enc, n_dec, H1=256, H2=128):
    """Inverse of :func:`flat_from_state_dict`; returns tensors on the flat buffer's device."""
    out, o = {}, 0

    def take(n, shape):
        nonlocal o
        t = flat[o:o + n].view(shape)
        o += n
        return t

    w1t = take(n_enc * H1, (n_enc, H1))
    out["encoder.blocks.0.weight"] = None  # placeholder to keep the reference's key order
    out["encoder.blocks.0.bias"] = None
    out["decoder.blocks.0.weight"] = None
    out["decoder.blocks.0.bias"] = None
    out["encoder_linear.weight"] = w1t.t().contiguous()
    out["encoder_linear.bias"] = take(H1, (H1,)).clone()
    out["encoder.blocks.0.weight"] = take(H2 * H1, (H2, H1)).clone()
    out["encoder.blocks.0.bias"] = take(H2, (H2,)).clone()
    out["decoder.blocks.0.weight"] = take(H1 * H2, (H1, H2)).clone()
    out["decoder.blocks.0.bias"] = take(H1, (H1,)).clone()
    out["decoder_linear.weight"] = take(n_dec * H1, (n_dec, H1)).clone()
    out["decoder_linear.bias"] = take(n_dec, (n_dec,)).clone()
    return out


class FastEpochLayout:
    """Production-mode layout (device-drawn dropout): the row order inside a batch is irrelevant, so the batch is the
    permutation slice itself — no per-batch sort on the host. Rows with neither data nor targets are still dropped
    and batches without data entries still skipped (src/models/ae.py:101, src/organization.py:153-155)."""

    def __init__(self, perm, batch_size, d_len, t_len, epoch_len=None):
        """epoch_len: ``perm`` is the concatenation of several epochs' permutations of that length (a whole round in one
        layout: batch ids keep counting across epochs, so one plan / one graph launch covers all of them)."""
        perm = np.asarray(perm, dtype=np.int64)
        n = len(perm)
        L = n if (epoch_len is None or epoch_len >= n) else int(epoch_len)
        n_ep = n // L if L else 0
        starts = (np.arange(0, L, batch_size)[None, :] + (np.arange(n_ep) * L)[:, None]).ravel() if n else \
            np.zeros(0, np.int64)
        nb = len(starts)
        dl, tl = d_len[perm], t_len[perm]
        keep = (dl + tl) > 0
        if n == 0:
            counts = np.zeros(0, np.int64)
            d_per = np.zeros(0, np.int64)
        else:
            # rows that are dropped have neither data nor targets: the per-batch sums can run over the unfiltered order
            d_per = np.add.reduceat(dl.astype(np.int64), starts)
            counts = np.add.reduceat(keep.astype(np.int64), starts)
        self.rows = perm if keep.all() else perm[keep]
        self.row_off = np.zeros(nb + 1, np.int32)
        self.row_off[1:] = np.cumsum(counts)
        self.active = (d_per > 0).tolist()
        self.batch_rows = counts.tolist()
        self.d_per_batch = d_per.tolist()
        self.n_t = int(tl.sum())
        self.n_d = int(dl.sum())


class RoundLayout:
    """All local epochs of one organization's round from ONE FastEpochLayout pass over the concatenated permutations
    (the per-epoch layouts of a round used to cost 360 numpy passes per 18-organization round, ~100 ms of host time
    next to a 200 ms GPU round: a slower host made the round host-bound). Per-epoch views are slices:
    rows [row_edges[e], row_edges[e+1]), local batch offsets off_local[e], entry counts n_t[e] / n_d[e]."""

    def __init__(self, perms, batch_size, d_len, t_len, n_rows, n_epochs):
        L = FastEpochLayout(perms, batch_size, d_len, t_len, epoch_len=n_rows)
        self.whole = L
        self.n_epochs = n_epochs
        self.nb_epoch = -(-n_rows // batch_size)                      # batches per epoch
        G = L.row_off.astype(np.int64)
        idx = np.arange(self.nb_epoch + 1)[None, :] + (np.arange(n_epochs) * self.nb_epoch)[:, None]
        self.row_edges = G[idx[:, 0]].tolist() + [int(G[-1])]          # first row of every epoch (+ end)
        self.off_local = (G[idx] - G[idx[:, :1]]).astype(np.int32)     # [n_epochs x (nb_epoch + 1)]
        self.rows = L.rows.astype(np.int32)
        self.off_global = L.row_off
        ct = np.concatenate([[0], np.cumsum(t_len[L.rows], dtype=np.int64)])
        cd = np.concatenate([[0], np.cumsum(d_len[L.rows], dtype=np.int64)])
        e = np.asarray(self.row_edges)
        self.n_t = (ct[e[1:]] - ct[e[:-1]]).tolist()
        self.n_d = (cd[e[1:]] - cd[e[:-1]]).tolist()
        self.n_t_total, self.n_d_total = L.n_t, L.n_d
        self.n_batches = len(L.active)


def decoder_default():
    """Decoder form of new engines: DMT_DECODER=gather|tc. Default gather (row-gather SDDMM + segmented reductions):
    measured on B200 at ML1M shape (500-row batches, 3.7 % dense targets) it is ~1.5x faster per step than the
    tcgen05 GEMMs with the 3xTF32 parity split; tc pays off for denser / larger batches."""
    mode = os.environ.get("DMT_DECODER", "gather")
    if mode not in ("tc", "gather"):
        raise ValueError("Not valid DMT_DECODER: {}".format(mode))
    return mode


class OrgEngine:
    """One organization's AAE on the device (Organization.train / predict, reference src/organization.py:140-217)."""

    def __init__(self, data: DeviceCSR, target: DeviceCSR, batch_rows, H1=256, H2=128, loss_kind=0, plan_epochs=1):
        self.data, self.target = data, target
        self.plan_epochs = plan_epochs
        self.n_rows, self.n_enc = data.shape
        self.n_dec = target.shape[1]
        self.H1, self.H2 = H1, H2
        self.batch_rows = batch_rows
        self.h = native.Org(self.n_rows, self.n_enc, self.n_dec, H1, H2, data.triple(), target.pair(), batch_rows,
                            loss_kind, plan_epochs=plan_epochs)
        self.d_len, self.t_len = data.row_len, target.row_len
        self.device = data.indptr.device
        self._keep_alive = []
        self.set_decoder(decoder_default() if target.sorted else "gather")

    def set_decoder(self, mode, passes=3):
        """'tc': decoder last layer as tcgen05 GEMMs (3xTF32 parity mode; passes=1 is reduced precision),
        'gather': SDDMM + segmented reductions."""
        if mode == "tc" and not self.target.sorted:
            raise ValueError("Not valid decoder mode: the tensor-core decoder needs sorted CSR column indices")
        self.decoder = mode
        self.h.set_decoder_mode(mode, passes)

    def set_round(self, params_flat, residual):
        """Fresh model + optimizer for the round (src/organization.py:144-148) and this round's targets."""
        self.h.wait_current()
        self.h.set_params(params_flat)
        self.h.set_target(residual)
        self._keep_alive = [params_flat, residual]

    def enqueue_epoch(self, layout: EpochLayout, keep=None, seed=0, hp=None, loss_out=None):
        rows = to_dev(layout.rows.astype(np.int32), self.device)
        off = to_dev(layout.row_off, self.device)
        if keep is not None and not keep.is_cuda:
            keep = to_dev(keep, self.device)
        self.h.wait_current()
        self.h.train_epoch(rows, off, layout.n_t, layout.n_d, keep=keep, seed=seed, epoch_loss=loss_out, **(hp or {}))
        self._keep_alive += [rows, off, keep, loss_out]

    def enqueue_epochs(self, layouts, seeds, hp=None, loss_out=None):
        """Several epochs with ONE host->device copy of all their row lists (device-generated dropout)."""
        nb = [len(l.row_off) - 1 for l in layouts]
        rows_all = to_dev(np.concatenate([l.rows for l in layouts]).astype(np.int32), self.device)
        off_all = to_dev(np.concatenate([l.row_off for l in layouts]), self.device)
        self.h.wait_current()
        r0 = o0 = l0 = 0
        for e, l in enumerate(layouts):
            rows = rows_all[r0:r0 + len(l.rows)]
            off = off_all[o0:o0 + nb[e] + 1]
            lo = loss_out[l0:l0 + nb[e]] if loss_out is not None else None
            self.h.train_epoch(rows, off, l.n_t, l.n_d, keep=None, seed=int(seeds[e]), epoch_loss=lo, **(hp or {}))
            r0 += len(l.rows)
            o0 += nb[e] + 1
            l0 += nb[e]
        self._keep_alive += [rows_all, off_all, loss_out]

    def enqueue_round(self, layout, seed, hp=None, loss_out=None):
        """All local epochs of a round as ONE plan and ONE graph launch (layout = FastEpochLayout with epoch_len;
        the engine must have been created with plan_epochs >= the number of epochs)."""
        rows = to_dev(layout.rows.astype(np.int32), self.device)
        off = to_dev(layout.row_off, self.device)
        self.h.wait_current()
        self.h.train_epoch(rows, off, layout.n_t, layout.n_d, keep=None, seed=int(seed), epoch_loss=loss_out,
                           **(hp or {}))
        self._keep_alive += [rows, off, loss_out]

    def params(self):
        return self.h.get_params()

    def predict(self, data: DeviceCSR, target: DeviceCSR, out):
        """Eval forward at the target positions of the rows the data matrix holds (a cold-start organization's train
        data is shorter than the global target: positions of the rows beyond it are left untouched)."""
        self.h.wait_current()
        n_rows = min(data.shape[0], target.shape[0])
        if self.decoder == "tc" and not target.sorted:  # this split's CSR cannot take the windowed epilogue
            self.h.set_decoder_mode("gather")
            self.h.predict(data.triple(), target.pair(), n_rows, out)
            self.h.set_decoder_mode("tc")
        else:
            self.h.predict(data.triple(), target.pair(), n_rows, out)
        return out

    def sync(self):
        self.h.sync()
        self._keep_alive = self._keep_alive[:2]

    def close(self):
        self.h.close()


class MtalState:
    """Global per-split state of the coordinator (Assist, reference src/assist.py:13-41) on the device."""

    def __init__(self, y: dict, data_split, target_mode, device="cuda", o_rows=None, org_row=None):
        """org_row: row of ``O_full`` that holds each organization's prediction vector (default: row = organization
        id); the sharded exchange uses a rank-blocked layout so that one in-place all-gather fills it (dist.py)."""
        self.splits = list(y)
        self.y = {k: DeviceCSR(y[k], device) for k in y}
        self.K = len(data_split)
        self.n_cols = y[self.splits[0]].shape[1]
        self.target_mode = target_mode
        self.loss_kind = native.LOSS_KIND[target_mode]
        self.device = device
        owner = np.full(self.n_cols, -1, np.int32)
        local = np.zeros(self.n_cols, np.int32)
        self.split_sizes = []
        for i, cols in enumerate(data_split):
            cols = np.asarray(cols, dtype=np.int64)
            owner[cols] = i
            local[cols] = np.arange(len(cols), dtype=np.int32)
            self.split_sizes.append(len(cols))
        if (owner < 0).any():
            raise ValueError("every column must belong to exactly one organization")
        self.owner_host, self.local_host = owner, local
        self.owner = torch.from_numpy(owner).to(device)
        self.data_split = [np.asarray(c, dtype=np.int64) for c in data_split]
        # O may carry padding rows so that an in-place all-gather over equal per-rank chunks can fill it (dist.py)
        self.O_full = {k: torch.zeros(max(self.K, o_rows or 0), self.y[k].nnz, device=device) for k in y}
        self.org_row_host = np.arange(self.K, dtype=np.int32) if org_row is None else np.asarray(org_row, np.int32)
        if len(self.org_row_host) != self.K or len(set(self.org_row_host.tolist())) != self.K:
            raise ValueError("org_row must give every organization its own row of O")
        self.org_row = None if org_row is None else torch.from_numpy(self.org_row_host).to(device)
        self._s_cold_cache = None
        self.O = {k: self.O_full[k][:self.K] for k in y} if org_row is None else None  # identity layout only
        self._views = {}
        self._eval_meta = {}
        self._combine_cache = None

    def O_orgmajor(self, split):
        """[K x nnz] copy in organization order (tests / inspection; the kernels index O_full through org_row)."""
        return self.O_full[split][torch.from_numpy(self.org_row_host.astype(np.int64)).to(self.device)]

    def o_row(self, split, org):
        """Organization ``org``'s prediction vector for ``split`` (a row view of O_full)."""
        return self.O_full[split][int(self.org_row_host[org])]

    def residual(self, F, split, clamp, out=None):
        return native.residual(F, self.y[split].data, self.loss_kind, 1.0 if clamp else 0.0, out)

    def evaluate(self, F, split="test", block_rows=500, topk=10, org=None):
        """Global metrics of the current prediction with the reference's test loop semantics
        (src/train_recsys_assist.py:175-217, src/logger.py:35-55) computed on the device: per-block Loss and
        RMSE (explicit) or NDCG@topk (implicit), entry-weighted over blocks. One launch + a [n_blocks x 3] read-back.
        org: score only that organization's columns (cold-start runs are scored on organization 0's,
        src/train_recsys_assist.py:180-182)."""
        y = self.y[split]
        n_rows = y.shape[0]
        key = (split, block_rows, topk, org)
        if key not in self._eval_meta:
            ip, ix = y.indptr_host, y.indices_host
            pos_dev = None
            if org is not None:
                pos = self.owner_view(split, org)["pos_host"]
                cnt = np.bincount(np.searchsorted(ip, pos, side="right") - 1, minlength=n_rows)[:n_rows]
                ip = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
                ix = ix[pos]
                pos_dev = torch.from_numpy(pos.astype(np.int64)).to(self.device)
            edges = np.arange(0, n_rows + block_rows, block_rows).clip(max=n_rows)
            m = (ip[edges[1:]] - ip[edges[:-1]]).astype(np.float64)          # entries per block
            rl = np.diff(ip) > 0
            rows_nz = np.add.reduceat(rl, edges[:-1]).astype(np.float64) if n_rows else np.zeros(0)
            k = np.array([min(topk, len(np.unique(ix[ip[a]:ip[b]]))) for a, b in zip(edges[:-1], edges[1:])],
                         dtype=np.int32)
            ip_dev = y.indptr if org is None else to_dev(ip.astype(np.int32), self.device)
            self._eval_meta[key] = (m, rows_nz, to_dev(k, self.device), ip_dev, pos_dev)
        m, rows_nz, k_dev, ip_dev, pos_dev = self._eval_meta[key]
        implicit = self.target_mode == "implicit"
        pred, tgt = (F, y.data) if pos_dev is None else (F[pos_dev].contiguous(), y.data[pos_dev].contiguous())
        sums = to_host(native.eval_blocks(ip_dev, pred, tgt, n_rows, block_rows, self.loss_kind,
                                          k_dev if implicit else None)).double().numpy()
        ok = m > 0
        w = m[ok] / m[ok].sum()
        out = {"{}/Loss".format(split): float((sums[ok, 0] / m[ok] * w).sum())}
        if implicit:
            out["{}/NDCG".format(split)] = float((sums[ok, 2] / rows_nz[ok] * w).sum())
        else:
            out["{}/RMSE".format(split)] = float((np.sqrt(sums[ok, 1] / m[ok]) * w).sum())
        return out

    def privatize(self, residual, mode, param, seed):
        """make_privacy (src/privacy.py) on the device, in place on a residual vector (production mode)."""
        native.privacy(residual, mode, param, seed, out=residual)
        return residual

    def owner_view(self, split, i):
        """Static per-owner view for the fit: positions of the owner's entries sorted by local column (stable),
        their rank in the original order, and the per-column segment offsets."""
        key = (split, i)
        if key not in self._views:
            ind = self.y[split].indices_host
            pos = np.flatnonzero(self.owner_host[ind] == i)
            idx = self.local_host[ind[pos]]
            order = np.argsort(idx, kind="stable")
            seg_off = np.zeros(self.split_sizes[i] + 1, np.int32)
            seg_off[1:] = np.cumsum(np.bincount(idx, minlength=self.split_sizes[i]))
            dev = self.device
            self._views[key] = {
                "pos_host": pos,
                "pos": torch.from_numpy(pos[order].astype(np.int32)).to(dev),
                "rank": torch.from_numpy(order.astype(np.int32)).to(dev),
                "seg_off": torch.from_numpy(seg_off).to(dev),
                "n": len(pos),
            }
        return self._views[key]

    def match_end(self, split, match_rate):
        """Global CSR position from which an owner's entries are UNmatched (partial alignment: only the first
        int(n * match_rate) entries of the owner's view see the other organizations, src/assist.py:95-103)."""
        ends = []
        for i in range(self.K):
            pos = self.owner_view(split, i)["pos_host"]
            n_match = int(len(pos) * match_rate)
            ends.append(pos[n_match] if n_match < len(pos) else self.y[split].nnz)
        return torch.tensor(ends, dtype=torch.int64, device=self.device)

    def fit_owner(self, i, F_prev, ar, ar_mode, aw_mode, match_rate, lr=0.1, steps=10, cold=False):
        """L-BFGS fit of owner i's assisted learning rates / assistance weights on the train split
        (src/assist.py:118-129). The two-loop recursion runs on tiny host vectors (torch.optim.LBFGS, as in the
        reference); every closure evaluation is ONE fused loss+gradient kernel over the owner's [n_i x K] view
        (cold start: the NaN-aware forward / backward pair of the differentiable module, dmt_assist_rows_*)."""
        n_rate = self.split_sizes[i]
        rate = torch.full((n_rate,), float(ar))
        weight = torch.ones(self.K) / self.K
        free = []
        if ar_mode == "optim":
            rate.requires_grad_(True)
            free.append(rate)
        if aw_mode == "optim":
            weight.requires_grad_(True)
            free.append(weight)
        if not free:
            return rate, weight
        v = self.owner_view("train", i)
        n = v["n"]
        n_match = int(n * match_rate) if match_rate < 1 else n
        h, t, V = native.assist_gather_view(F_prev, self.y["train"].data, self.O_full["train"], v["pos"], v["rank"], i,
                                            n_match, org_row=self.org_row, K=self.K)
        scratch = torch.empty(native.load().dmt_assist_scratch_floats(self.K), device=self.device)
        opt = torch.optim.LBFGS(free, lr=lr)
        # all closure inputs / outputs of one evaluation travel in ONE pinned staging pair (no pageable copies, one sync)
        stage_in = torch.empty(n_rate + self.K, dtype=torch.float32, pin_memory=True)
        dev_in = torch.empty(n_rate + self.K, device=self.device)
        if cold:
            if "idx" not in v:
                seg = v["seg_off"].cpu().numpy()
                v["idx"] = torch.from_numpy(np.repeat(np.arange(n_rate, dtype=np.int32), np.diff(seg))).to(self.device)
                v["seg"] = native.sort_segments(v["idx"], n_rate)

        def closure():
            stage_in[:n_rate].copy_(rate.detach())
            stage_in[n_rate:].copy_(weight.detach())
            dev_in.copy_(stage_in, non_blocking=True)
            r_d, w_d = dev_in[:n_rate], dev_in[n_rate:]
            if cold:
                tgt, q = native.assist_rows_fwd(V, 1, n, h, v["idx"], r_d, w_d, n, self.K)
                dpred, sums = native.loss_fwd(tgt, t, self.loss_kind, True)
                d_rate, d_w = native.assist_rows_bwd(V, 1, n, v["idx"], r_d, w_d, q, dpred, v["seg"], n, self.K)
                out = to_host(torch.cat([sums[:1], d_rate, d_w])) / n
            else:
                out = to_host(native.assist_loss_grad(h, t, V, v["seg_off"], r_d, w_d, self.loss_kind, scratch))
            if rate.requires_grad:
                rate.grad = out[1:1 + n_rate].clone()
            if weight.requires_grad:
                weight.grad = out[1 + n_rate:].clone()
            return out[0]

        for _ in range(steps):
            opt.step(closure)
        return rate.detach(), weight.detach()

    def fit_owners_device(self, F_prev, ar, ar_mode, aw_mode, match_rate, lr=0.1, steps=10):
        """All owners' L-BFGS fits (src/assist.py:118-129) with the optimizer itself on the device
        (dmt_assist_fit: closure, two-loop recursion, step and torch's stopping rules as a chain of launches, no host
        round trip), the owners spread over a few side streams; ONE device->host copy of all fitted rates / weights at
        the end. (The host-driven variant, fit_owner, pays one synchronous read-back per closure evaluation and
        torch.optim.LBFGS's per-iteration Python walk over up to 100 history pairs: ~110 ms per Amazon-shape round.)"""
        main = torch.cuda.current_stream()
        n_side = min(self.K, 4)
        if getattr(self, "_fit_streams", None) is None or len(self._fit_streams) < n_side:
            self._fit_streams = [torch.cuda.Stream(device=self.device) for _ in range(n_side)]
        sizes = [self.split_sizes[i] + self.K for i in range(self.K)]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        init = np.empty(int(offs[-1]), np.float32)
        for i in range(self.K):
            init[offs[i]:offs[i] + self.split_sizes[i]] = float(ar)
            init[offs[i] + self.split_sizes[i]:offs[i + 1]] = 1.0 / self.K
        params = to_dev(init, self.device)
        ready = main.record_event()
        keep = []
        for i in range(self.K):
            v = self.owner_view("train", i)
            n = v["n"]
            n_match = int(n * match_rate) if match_rate < 1 else n
            side = self._fit_streams[i % n_side]
            side.wait_event(ready)
            with torch.cuda.stream(side):
                h, t, V = native.assist_gather_view(F_prev, self.y["train"].data, self.O_full["train"], v["pos"],
                                                    v["rank"], i, n_match, org_row=self.org_row, K=self.K)
                keep.append((h, t, V) + native.assist_fit(h, t, V, v["seg_off"], params[offs[i]:offs[i + 1]],
                                                          ar_mode == "optim", aw_mode == "optim", self.loss_kind,
                                                          lr=lr, steps=steps))
        for side in self._fit_streams[:n_side]:
            main.wait_stream(side)
        host = to_host(params)  # synchronises the main stream: every buffer in `keep` has been consumed
        del keep
        fitted = []
        for i in range(self.K):
            r = host[offs[i]:offs[i] + self.split_sizes[i]].clone()
            w = host[offs[i] + self.split_sizes[i]:offs[i + 1]].clone()
            fitted.append((r, w))
        return fitted

    def update(self, F_prev: dict, ar, ar_mode="constant", aw_mode="constant", match_rate=1.0, out=None, cold=False):
        """Assist.update (src/assist.py:81-179): fit per owner on train, then ONE combine pass per split.
        cold: cold-start run — organization 0's output is NaN on the aligned rows it never saw; those entries combine
        organizations 1.. with softmax(w[1:]) (src/models/assist.py:28-34)."""
        if (ar_mode == "optim" or aw_mode == "optim") and not cold and os.environ.get("DMT_LBFGS", "device") != "host":
            fitted = self.fit_owners_device(F_prev["train"], ar, ar_mode, aw_mode, match_rate)
        else:
            fitted = [self.fit_owner(i, F_prev["train"], ar, ar_mode, aw_mode, match_rate, cold=cold)
                      for i in range(self.K)]
        rate_col = np.zeros(self.n_cols, np.float32)
        S = np.zeros((self.K, self.K), np.float32)
        S_cold = np.zeros((self.K, self.K), np.float32) if cold else None
        for i, (rate, weight) in enumerate(fitted):
            rate_col[self.data_split[i]] = rate.numpy()
            S[i] = torch.softmax(weight, -1).numpy()
            if cold and self.K > 1:
                S_cold[i, 1:] = torch.softmax(weight[1:], -1).numpy()
        # constant rates / weights repeat every round: upload once (a pageable copy on the compute stream would make the
        # host wait for the whole round that the stream is still ordered behind)
        ck = (rate_col.tobytes(), S.tobytes(), cold)
        if self._combine_cache is None or self._combine_cache[0] != ck:
            self._combine_cache = (ck, to_dev(rate_col, self.device), to_dev(S, self.device),
                                   to_dev(S_cold, self.device) if cold else None)
        rate_col_d, S_d, S_cold_d = self._combine_cache[1], self._combine_cache[2], self._combine_cache[3]
        F_next = {}
        for k in self.splits:
            me = self.match_end(k, match_rate) if match_rate < 1 else None
            F_next[k] = native.assist_combine(F_prev[k], self.O_full[k], self.y[k].indices, self.owner, rate_col_d, S_d,
                                              me, out[k] if out is not None else None, org_row=self.org_row,
                                              S_cold=S_cold_d, K=self.K)
        return F_next, fitted


def he_seed(seed, *parts):
    """Deterministic 64-bit seed for the on-device dropout generator."""
    x = (int(seed) * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
    for p in parts:
        x = ((x ^ (int(p) + 0x632BE59BD9B4E019)) * 0xD6E8FEB86659FD93) & (2 ** 64 - 1)
    return x


def bytes_per_round(K, nnz_train, nnz_test, n_rows, n_enc_list, n_dec, epochs, batch_rows, H1=256, H2=128):
    """Algorithmic (compulsory) bytes of one assist round, per SURVEY.md §8d: decoder SDDMM 4*H1+16 B per target
    visit forward+first-backward, segmented dW4 4*H1+8 per target, dense Adam 28 B/param/step, residual 12 B and
    combine (4K+12) B per rating."""
    steps = epochs * math.ceil(n_rows / batch_rows)
    total = 0
    for n_enc in n_enc_list:
        n_params = n_enc * H1 + H1 + H2 * H1 + H2 + H1 * H2 + H1 + n_dec * H1 + n_dec
        total += epochs * nnz_train * ((4 * H1 + 16) + (4 * H1 + 8))  # decoder fwd/dZ3 + dW4 reduction
        total += steps * n_params * (28 + 8)  # Adam + gradient zero/norm pass
        total += (nnz_train + nnz_test) * (4 * H1 + 12)  # predict
    total += (nnz_train + nnz_test) * (12 + 4 * K + 12)
    return total
