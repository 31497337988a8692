"""Device-resident assistance rounds with organizations sharded over ranks.

One process per GPU. Every rank keeps the global CSR structure, ground truth and current global prediction F_t;
organization k (its data column block, parameters, optimizer state, epoch plans and CUDA graph) lives on the rank
``dist.assign_orgs`` gives it (balanced by count, then by work). Inside a round the ranks share nothing; the only
exchange is the K prediction vectors after ``predict`` (reference: plain Python list passing,
src/train_recsys_assist.py:166-172; src/assist.py:81-84): an in-place NCCL all-gather of the rank-blocked
organization-major matrix rows over NVLink. After it every
rank runs the same (cheap, O(K nnz)) ``update`` so F_{t} is replicated without a second collective.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import engine as E
from . import native


# organizations per GPU up to which a step's backward pass is enqueued as parallel graph branches (dmt_org_set_fanout);
# DMT_FANOUT=0|1 overrides. Measured on one B200 at ML1M shape, ms per round without -> with: 3 organizations 72.9 -> 63.3,
# 5: 96.1 -> 85.5, 9: 140.4 -> 130.7; at 18 the round is throughput-bound and the shape of the step does not matter
# (214.7 vs 215.0 once the host no longer stalls behind the compute streams).
FANOUT_MAX_ORGS = 9
DEC_BLOCKS_MANY_ORGS = 111  # decoder chunk grid when a rank holds more organizations than that (default 296 = 2 per SM)


def decoder_blocks_for(n_orgs):
    """Grid of the decoder chunk kernel by the number of organizations that share the GPU (0 = the kernel's default,
    two blocks per SM). Measured at ML1M shape, ms per round: 18 organizations 296 -> 214.7, 148 -> 205.5, 111 -> 204.3;
    9 organizations 296 -> 106.8, 222 -> 105.6, 148 -> 104.4."""
    if n_orgs > FANOUT_MAX_ORGS:
        return DEC_BLOCKS_MANY_ORGS
    if n_orgs >= 6:
        return 148
    return 0


def row_tile_for(n_orgs):
    """Batch rows per CTA of the fused step's row kernels: always 8. Four rows per CTA (two warps per encoder row)
    shorten a lone organization's step (19 + 16 -> 14 + 12 us) and the round of a rank with one or two organizations
    (30.5 -> 28.6 ms; three: 35.8 vs 36.0; five and more: slower), but the tile size changes the grouping of the
    encoder and bias-gradient sums, so choosing it by the organizations a rank happens to hold would make the result
    depend on the sharding in the last bit (scripts/check_sharded_dropin.py caught exactly that). An 8-GPU run is
    bounded by its three-organization ranks anyway. `dmt_org_set_row_tile` / DMT_STREAM_ROWS=4 remain for explicit use."""
    return 8


# device memory the whole-round plans of a rank may take (bytes); DMT_WHOLE_ROUND=0|1 overrides
WHOLE_ROUND_PLAN_BUDGET = 48 << 30
# Measured on one B200 at ML1M shape (ms per round, per-epoch -> whole-round): 3 organizations 63.0 -> 59.9, 9
# organizations 111.2 -> 108.6, 18 organizations 214.7 -> 214.0 (throughput-bound: no difference), so it is used for
# ranks that hold few organizations, where the per-organization chain of dependent steps is what bounds a round.
WHOLE_ROUND_MAX_ORGS = 9


def xavier_uniform_(shape, device, generator=None):
    """nn.init.xavier_uniform_ on the device (production mode: same distribution as the reference's init,
    src/models/ae.py:22-28,89-96, drawn by the device generator instead of torch's CPU generator)."""
    fan_out, fan_in = shape
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return torch.empty(shape, device=device).uniform_(-a, a, generator=generator)


def init_flat_params(n_enc, n_dec, H1, H2, device, generator=None):
    """Fresh AAE parameters in the engine's flat layout (W1t b1 W2 b2 W3 b3 W4 b4), xavier weights, zero biases."""
    W1 = xavier_uniform_((H1, n_enc), device, generator)
    W2 = xavier_uniform_((H2, H1), device, generator)
    W3 = xavier_uniform_((H1, H2), device, generator)
    W4 = xavier_uniform_((n_dec, H1), device, generator)
    z = lambda n: torch.zeros(n, device=device)
    return torch.cat([W1.t().contiguous().view(-1), z(H1), W2.view(-1), z(H2), W3.view(-1), z(H1), W4.view(-1),
                      z(n_dec)]).contiguous()


class AssistRounds:
    def __init__(self, mats, data_split, target_mode, batch_rows, clamp=False, ar=0.1, ar_mode="constant",
                 aw_mode="constant", match_rate=1.0, local_epochs=20, rank=0, world=1, device="cuda", H1=256, H2=128,
                 seed=0, hp=None, group=False, privacy=None, whole_round=None):
        """mats: {'train': (data_csr, target_csr), 'test': (...)} global scipy CSR matrices (rows = aligned entity)."""
        self.rank, self.world, self.device = rank, world, device
        self.K = len(data_split)
        self.target_mode, self.clamp = target_mode, clamp
        self.ar, self.ar_mode, self.aw_mode, self.match_rate = ar, ar_mode, aw_mode, match_rate
        self.local_epochs, self.batch_rows, self.seed = local_epochs, batch_rows, seed
        self.privacy = privacy  # None or (mode 'dp'|'ip', param): device-side make_privacy on the broadcast residuals
        self.H1, self.H2 = H1, H2
        self.hp = hp or dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=5e-4, max_norm=1.0)
        self.splits = list(mats)
        cols = [np.asarray(c, dtype=np.int64) for c in data_split]
        y = {k: mats[k][1] for k in mats}
        for m in y.values():
            m.sort_indices()
        # organizations -> ranks, balanced by count then by work (target entries are the same for every organization;
        # the data entries and the encoder's share of the parameters differ); rank r's rows of O_full are
        # [r*chunk, (r+1)*chunk) so ONE in-place all-gather per split publishes them (dist.py)
        from . import dist as D
        d_nnz = np.diff(mats["train"][0].tocsc().indptr)
        costs = [float(y["train"].nnz + 2 * d_nnz[c].sum()) for c in cols]
        owned, self.chunk, org_row = D.assign_orgs(costs, world)
        self.state = E.MtalState(y, cols, target_mode, device, o_rows=self.chunk * world,
                                 org_row=org_row if world > 1 else None)
        self.n_rows = y["train"].shape[0]
        self.my_orgs = owned[rank]
        self.org_data, self.org_test_data, self.eng = {}, {}, {}
        # One plan + one graph per organization and ROUND (instead of per local epoch) when the plan buffers fit: ~40 B
        # per target entry and planned epoch, i.e. 0.7 GB per organization at ML1M shape with 20 epochs. Same batches,
        # same step order, same dropout draws -> identical results; 20x fewer plan kernels and graph launches.
        nnz_t = y["train"].nnz
        plan_bytes = 40 * nnz_t * local_epochs * max(1, len(self.my_orgs))
        env = os.environ.get("DMT_WHOLE_ROUND")
        if whole_round is None:
            whole_round = (plan_bytes < WHOLE_ROUND_PLAN_BUDGET and len(self.my_orgs) <= WHOLE_ROUND_MAX_ORGS) \
                if env is None else env == "1"
        # 32-bit (batch, column) sort keys and int32 entry offsets bound what one plan can cover
        n_b = -(-self.n_rows // batch_rows) * local_epochs
        if n_b * max(y["train"].shape[1], 1) >= 2 ** 32 or nnz_t * local_epochs >= 2 ** 31 - 2:
            whole_round = False
        self.whole_round = bool(whole_round) and not group
        for k in self.my_orgs:
            d = E.DeviceCSR(mats["train"][0][:, cols[k]].tocsr(), device)
            self.org_data[k] = d
            # the test split's data is the train matrix (reference src/datasets/movielens.py:367-371)
            same = mats["test"][0] is mats["train"][0]
            self.org_test_data[k] = d if same else E.DeviceCSR(mats["test"][0][:, cols[k]].tocsr(), device)
            self.eng[k] = E.OrgEngine(d, self.state.y["train"], batch_rows, H1, H2, native.LOSS_KIND[target_mode],
                                      plan_epochs=local_epochs if (group or self.whole_round) else 1)
        # optional lockstep group: one launch per step kernel for ALL organizations of this rank (dmt_group_*).
        # Measured on B200 at ML1M shape (18 organizations): 281 ms/round against 257 ms for per-organization graphs
        # on private streams — both are bound by the same L2 gather traffic — so per-organization graphs stay default.
        fan = os.environ.get("DMT_FANOUT")
        fan = (len(self.my_orgs) <= FANOUT_MAX_ORGS) if fan is None else fan == "1"
        # many organizations per GPU: a smaller grid for the register-heavy decoder chunk kernel leaves room for the
        # other organizations' kernels (18 organizations: 214.7 ms per round with 296 blocks, 204.3 with 111, 203.9 with 74)
        dec_blocks = os.environ.get("DMT_DEC_BLOCKS")
        dec_blocks = decoder_blocks_for(len(self.my_orgs)) if dec_blocks is None else int(dec_blocks)
        for k in self.my_orgs:
            self.eng[k].h.set_fanout(fan)
            self.eng[k].h.set_decoder_blocks(dec_blocks)
            self.eng[k].h.set_row_tile(row_tile_for(len(self.my_orgs)))
        self.fanout = fan
        self.group = native.Group([self.eng[k].h for k in self.my_orgs]) if group and self.my_orgs else None
        self.cols = cols
        self.mats = mats
        self.residual = {k: torch.empty(self.state.y[k].nnz, device=device) for k in self.splits}
        self.F = None
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed * 1000003 + rank)
        self.host_gen = torch.Generator()
        self.host_gen.manual_seed(seed * 7919 + rank)
        self.round_losses = {}
        self._up_stream = None
        self._pool = None
        self._pending_uploads = []
        self._held_uploads = []

    # ------------------------------------------------------------------ round 0 (replicated: cheap)
    def round0(self):
        """models.base per organization -> F_0 (src/organization.py:29-138); every rank computes all organizations."""
        F = {k: torch.empty(self.state.y[k].nnz, device=self.device) for k in self.splits}
        implicit = self.target_mode == "implicit"
        bs = self.batch_rows
        train_data = self.mats["train"][0]
        tr = E.DeviceCSR(train_data, self.device)
        # per-column sums over the whole train data == per-organization base on its own column block
        base = torch.zeros(tr.shape[1], device=self.device)
        count = torch.zeros(tr.shape[1], device=self.device)
        native.base_fit(tr.indices, tr.data, base, count)
        owner = self.state.owner_host
        for k in self.splits:
            yk = self.state.y[k]
            if not implicit:
                # unseen columns take the mean of the seen means OF THEIR OWN organization
                pred = torch.empty(yk.nnz, device=self.device)
                for i in range(self.K):
                    ci = torch.from_numpy(self.cols[i]).to(self.device)
                    pos = torch.from_numpy(self.state.owner_view(k, i)["pos_host"]).to(self.device)
                    local = torch.from_numpy(self.state.local_host[yk.indices_host[self.state.owner_view(k, i)["pos_host"]]]
                                             ).to(self.device)
                    pred[pos] = native.base_predict(base[ci].contiguous(), count[ci].contiguous(), local.contiguous(),
                                                    False)
                F[k] = pred
            else:
                # count = sum over the organization's loader batches of #rows with data (src/models/base.py:35-37)
                imp = np.zeros(self.K, np.float32)
                for i in range(self.K):
                    rl = np.diff(train_data[:, self.cols[i]].tocsr().indptr)
                    imp[i] = sum(int((rl[s:s + bs] > 0).sum()) for s in range(0, len(rl), bs))
                denom = torch.from_numpy(imp[owner]).to(self.device)
                F[k] = base[yk.indices.long()] / denom[yk.indices.long()]
        self.F = F
        return F

    # ------------------------------------------------------------------ one assistance round
    def run_round(self, t, exchange=None):
        """make_dataset -> local training of this rank's organizations -> predict -> exchange -> update."""
        self.train_predict(t)
        if exchange is not None:
            exchange(self.state.O)
        return self.combine()

    def train_predict(self, t):
        """Rank-local part of a round. Parameter init, sampler permutations and the dropout stream are seeded per
        (seed, organization, round), so the result does not depend on how organizations are sharded over ranks."""
        st = self.state
        for i, k in enumerate(self.splits):
            st.residual(self.F[k], k, self.clamp, out=self.residual[k])
            if self.privacy is not None:  # src/assist.py:59-60; same noise on every rank (seeded by round and split)
                st.privatize(self.residual[k], self.privacy[0], self.privacy[1], E.he_seed(self.seed, t, i, 1 << 22))
        loss_bufs = {}
        layouts, rows_dev, off_dev, glob_dev = {}, {}, {}, {}
        # (1) host-only preparation of every organization's round: no CUDA call, so it overlaps the previous round that
        #     the GPU is still executing (run_round never synchronises). One layout pass per organization and round,
        #     organizations spread over a few host threads (numpy and torch.randperm release the GIL).
        need_glob = self.whole_round or self.group is not None

        def prepare(org):
            eng = self.eng[org]
            g = torch.Generator()
            g.manual_seed(E.he_seed(self.seed, org, t, 1 << 21) & (2 ** 63 - 1))
            perms = np.concatenate([torch.randperm(self.n_rows, generator=g).numpy() for _ in range(self.local_epochs)])
            return E.RoundLayout(perms, self.batch_rows, eng.d_len, eng.t_len, self.n_rows, self.local_epochs)

        if len(self.my_orgs) > 1:
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=min(8, len(self.my_orgs)))
            layouts = dict(zip(self.my_orgs, self._pool.map(prepare, self.my_orgs)))
        else:
            layouts = {org: prepare(org) for org in self.my_orgs}
        # (2) uploads from pinned staging on a side stream: not ordered behind the compute streams, so neither the
        #     copies nor the host wait for the previous round to drain
        self._reap_uploads()
        up = self._upload_stream()
        held = []
        with torch.cuda.stream(up):
            for org in self.my_orgs:
                lay = layouts[org]
                rows_dev[org] = self._upload(lay.rows)
                held.append(rows_dev[org])
                if need_glob:
                    glob_dev[org] = self._upload(lay.off_global)
                    held.append(glob_dev[org])
                else:
                    off_dev[org] = self._upload(lay.off_local.ravel())
                    held.append(off_dev[org])
        uploaded = up.record_event()
        torch.cuda.current_stream().wait_event(uploaded)
        # (3) device side: fresh parameters, targets, loss buffers
        for org in self.my_orgs:
            eng = self.eng[org]
            self.gen.manual_seed(E.he_seed(self.seed, org, t, 1 << 20) & (2 ** 63 - 1))
            flat0 = init_flat_params(eng.n_enc, eng.n_dec, self.H1, self.H2, self.device, self.gen)
            eng.set_round(flat0, self.residual["train"])
            loss_bufs[org] = torch.zeros(layouts[org].n_batches, device=self.device)
            eng.h.wait_current()
            eng._keep_alive += [loss_bufs[org]]
        self._held_uploads = held
        same_rows = len({len(rows_dev[o]) for o in self.my_orgs}) == 1
        if self.group is not None and same_rows:
            # whole round of every organization: one plan per organization + ONE graph launch for all steps
            lays = [layouts[o] for o in self.my_orgs]
            self.group.train([rows_dev[o] for o in self.my_orgs], [glob_dev[o] for o in self.my_orgs],
                             len(rows_dev[self.my_orgs[0]]), lays[0].n_batches, [l.n_t_total for l in lays],
                             [l.n_d_total for l in lays], [E.he_seed(self.seed, o, t, 0) for o in self.my_orgs],
                             batch_loss=[loss_bufs[o] for o in self.my_orgs], **self.hp)
            self._finish_round_local(t, loss_bufs)
            return
        if self.whole_round:
            # one plan + one graph launch per organization for all local epochs of the round
            for org in self.my_orgs:
                lay = layouts[org]
                self.eng[org].h.train_epoch(rows_dev[org], glob_dev[org], lay.n_t_total, lay.n_d_total, keep=None,
                                            seed=E.he_seed(self.seed, org, t, 0), epoch_loss=loss_bufs[org], **self.hp)
            self._finish_round_local(t, loss_bufs)
            return
        # per-organization graphs, epoch-major enqueue: every organization's stream gets work early, so the GPU never
        # waits for the host to reach the last organization
        for e in range(self.local_epochs):
            for org in self.my_orgs:
                lay = layouts[org]
                nb = lay.nb_epoch
                self.eng[org].h.train_epoch(rows_dev[org][lay.row_edges[e]:lay.row_edges[e + 1]],
                                            off_dev[org][e * (nb + 1):(e + 1) * (nb + 1)], lay.n_t[e], lay.n_d[e],
                                            keep=None, seed=E.he_seed(self.seed, org, t, 0),  # the step counter varies the draw
                                            epoch_loss=loss_bufs[org][e * nb:(e + 1) * nb], **self.hp)
        self._finish_round_local(t, loss_bufs)

    def _finish_round_local(self, t, loss_bufs):
        st = self.state
        for org in self.my_orgs:
            eng = self.eng[org]
            eng.predict(self.org_data[org], st.y["train"], st.o_row("train", org))
            eng.predict(self.org_test_data[org], st.y["test"], st.o_row("test", org))
        for org in self.my_orgs:
            self.eng[org].h.signal_current()  # the current stream waits for every organization's stream
        self.round_losses[t] = loss_bufs
        # the uploaded row lists live on the side stream's pool: keep them referenced until this round has run
        self._pending_uploads.append((torch.cuda.current_stream().record_event(), self._held_uploads))
        self._held_uploads = []

    def _upload_stream(self):
        if self._up_stream is None:
            self._up_stream = torch.cuda.Stream(device=self.device)
        return self._up_stream

    def _upload(self, arr):
        """numpy -> pinned staging (torch's caching host allocator: same sizes every round, so no cudaHostAlloc after
        the first rounds) -> device, asynchronously on the current (upload) stream."""
        t = torch.from_numpy(arr)
        E.XFER["h2d"] += t.numel() * t.element_size()
        pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pin.copy_(t)
        return pin.to(self.device, non_blocking=True)

    def _reap_uploads(self):
        self._pending_uploads = [(ev, ts) for ev, ts in self._pending_uploads if not ev.query()]

    def combine(self):
        """Replicated part: Assist.update over the exchanged outputs -> F_t on every rank."""
        F_next, fitted = self.state.update(self.F, self.ar, self.ar_mode, self.aw_mode, self.match_rate)
        self.F = F_next
        return F_next, fitted

    def evaluate(self, split="test"):
        """Test metrics of the current global prediction F_t, computed on the device (MtalState.evaluate)."""
        return self.state.evaluate(self.F[split], split, self.batch_rows)

    def sync(self):
        for eng in self.eng.values():
            eng.sync()
        torch.cuda.synchronize()

    def rating_visits_per_round(self):
        """Unit of work: K * (epochs * nnz_train + nnz_train + nnz_test)."""
        n_tr, n_te = self.state.y["train"].nnz, self.state.y["test"].nnz
        return self.K * (self.local_epochs * n_tr + n_tr + n_te)

    def close(self):
        if self.group is not None:
            self.group.close()
            self.group = None
        for eng in self.eng.values():
            eng.close()
