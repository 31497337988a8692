"""Drop-in ``organization`` module: ``Organization`` with the reference's surface
(src/organization.py:21-217: ``initialize`` / ``train`` / ``predict`` and the attributes ``organization_id``,
``data_split``, ``num_items``, ``model_name``, ``model_state_dict``), running on the device-resident engine.

Inputs are the reference's dataset objects (``.data`` / ``.target`` scipy CSR in host memory, ``.num_users``,
``.num_items``); outputs are fresh host objects (scipy CSR fp32, CPU state_dicts). Device buffers are owned here and
reused across calls, keyed by the content of the CSR structure.
"""
import sys
import weakref

import os

import numpy as np
import torch
from scipy.sparse import csr_matrix

from dmtcdr_b200 import engine as E
from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg

import models

_DEVICE_CSR = {}
_CSR_IDENTITY = {}  # (id(indptr), id(indices)) -> (indptr, indices, shape, nnz, content key)


def _structure_key(m):
    """Content key of the CSR structure. CRC-ing the index arrays costs ~1 ms per call at ML1M shape and the drivers
    hand the same arrays back ~100 times per round, so the arrays' identity is tried first; the cache holds references
    to them, which keeps their ids unique for as long as an entry lives."""
    ip, ix = m.indptr, m.indices
    fk = (id(ip), id(ix))
    ent = _CSR_IDENTITY.get(fk)
    if ent is not None and ent[0] is ip and ent[1] is ix and ent[2] == m.shape and ent[3] == m.nnz:
        return ent[4]
    key = E.csr_key(m)
    if len(_CSR_IDENTITY) > 512:
        _CSR_IDENTITY.clear()
    _CSR_IDENTITY[fk] = (ip, ix, m.shape, m.nnz, key)
    return key


def device_csr(m, with_values=True):
    """Device copy of a scipy CSR, reused while the structure (and, if requested, the values) are unchanged."""
    key = _structure_key(m)
    hit = _DEVICE_CSR.get(key)
    if hit is None:
        hit = E.DeviceCSR(m, _device(), with_values=False)
        hit._values_crc = None
        _DEVICE_CSR[key] = hit
    if with_values:
        host = np.ascontiguousarray(m.data, dtype=np.float32)
        crc = E.zlib.crc32(host.view(np.uint8))
        if hit._values_crc != crc:
            hit.data = E.to_dev(host, _device())
            hit._values_crc = crc
    return hit


_D2H_STREAMS = {}


def _d2h_stream(dev):
    st = _D2H_STREAMS.get(dev)
    if st is None:
        st = _D2H_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return st


def _device():
    dev = str(cfg['device'])
    if not dev.startswith('cuda'):
        raise native.NativeError("dmtcdr_b200 has no CPU path: run with --device cuda (got '{}')".format(dev))
    return dev


def _rng_mode():
    """'reference': consume torch's CPU generator exactly like the reference (data-loader seeds, parameter init,
    dropout masks drawn on the host) so a run replays bit-for-bit comparable batches; 'device' (default): same
    parameter init and sampler, dropout drawn by the on-device counter-based generator."""
    if 'dmt_rng' in cfg:
        return cfg['dmt_rng']
    return os.environ.get('DMT_RNG', 'device')  # an unmodified driver cannot add cfg keys: the environment can


def _shape_target():
    if cfg['data_mode'] == 'user':
        return (cfg['num_users']['target'], cfg['num_items']['target'])
    if cfg['data_mode'] == 'item':
        return (cfg['num_items']['target'], cfg['num_users']['target'])
    raise ValueError('Not valid data mode')


# Work the drop-in defers so that the host never waits inside the per-organization API calls of a round: train-loss
# logging callbacks and the host copies of predictions. `flush_pending()` runs at the next natural barrier
# (Assist.update, which needs every organization's output anyway) or when somebody actually reads the data.
_PENDING = []


def flush_pending():
    global _PENDING
    todo, _PENDING = _PENDING, []
    for cb in todo:
        cb()


class LazyCSR(csr_matrix):
    """scipy CSR whose ``data`` is still in flight from the device: the device->host copy into a pinned buffer is
    enqueued behind the organization's work, and the values are moved into this matrix' own (fresh, pageable) array the
    first time ``data`` is read. The reference's driver hands predict() outputs straight to Assist.update, which takes
    the device-resident copy (``_dmt_pred_dev``) — in that flow the host copy is never waited for."""

    @property
    def data(self):
        f = self.__dict__.get('_dmt_force')
        if f is not None:
            self.__dict__['_dmt_force'] = None
            f()
        return self.__dict__['_dmt_data']

    @data.setter
    def data(self, v):
        self.__dict__['_dmt_data'] = v

    def __reduce__(self):  # pickles / deep-copies as a plain CSR
        return (csr_matrix, ((self.data, self.indices, self.indptr), self.shape))


def lazy_csr(vec_dev, indices_host, indptr_host, shape, slot, wait_stream=None):
    """A LazyCSR over `vec_dev` (device fp32 vector in the CSR's storage order): the device->host copy is enqueued on the
    side stream behind `wait_stream` (default: the current stream) into `slot['pin']`, a persistent pinned buffer that
    consecutive outputs of the same size share (an earlier, still unread output gets its values first); nobody waits."""
    dev = vec_dev.device
    n = vec_dev.numel()
    if slot.get('pin') is None or slot['pin'].numel() != n:
        slot['pin'] = torch.empty(n, dtype=torch.float32, pin_memory=True)
        slot['last'] = None
    prev = slot['last']() if slot.get('last') is not None else None
    if prev is not None:
        prev.data
    src = wait_stream if wait_stream is not None else torch.cuda.current_stream(dev)
    ready = src.record_event()
    d2h = _d2h_stream(str(dev))
    with torch.cuda.stream(d2h):
        d2h.wait_event(ready)
        slot['pin'].copy_(vec_dev, non_blocking=True)
        done = d2h.record_event()
    vec_dev.record_stream(d2h)
    E.XFER["d2h"] += n * 4
    host = np.empty(n, dtype=np.float32)
    m = LazyCSR((host, indices_host.astype(np.int32, copy=False), indptr_host.astype(np.int32, copy=False)),
                shape=shape, copy=False)
    pin = slot['pin']
    dest = m.__dict__['_dmt_data']  # the array scipy actually kept

    def force():
        done.synchronize()
        np.copyto(dest, pin.numpy())

    m.__dict__['_dmt_force'] = force
    slot['last'] = weakref.ref(m)
    return m


class LazyStateDict(dict):
    """state_dict whose tensors are still on the organization's stream; materialised on first read."""

    def __init__(self, flat, eng, n_enc, n_dec, H1, H2, on_ready=None):
        super().__init__()
        self._pending = (flat, eng, n_enc, n_dec, H1, H2, on_ready)
        self._on_ready = on_ready
        self._eng = eng

    def _log(self):
        """Run the train-loss logging callback once (needs the organization's training to have finished)."""
        cb = self.__dict__.get('_on_ready')
        if cb is not None:
            self.__dict__['_on_ready'] = None
            eng = self.__dict__.get('_eng')
            if eng is not None:
                eng.h.sync()  # the losses are written on the organization's stream
            cb()

    def _force(self):
        p = self.__dict__.get('_pending')
        if p is not None:
            self.__dict__['_pending'] = None
            flat, eng, n_enc, n_dec, H1, H2, on_ready = p
            eng.sync()
            sd = E.state_dict_from_flat(flat, n_enc, n_dec, H1, H2)
            dict.update(self, {k: E.to_host(v) for k, v in sd.items()})
            self._log()
        return self

    def flat_device(self):
        p = self.__dict__.get('_pending')
        return p[0] if p is not None else None

    def __getitem__(self, k):
        return dict.__getitem__(self._force(), k)

    def __iter__(self):
        return dict.__iter__(self._force())

    def __len__(self):
        return dict.__len__(self._force())

    def keys(self):
        return dict.keys(self._force())

    def items(self):
        return dict.items(self._force())

    def values(self):
        return dict.values(self._force())

    def __contains__(self, k):
        return dict.__contains__(self._force(), k)

    def __reduce__(self):
        return (dict, (dict(self._force()),))


_LAYOUTS = {}
_LAYOUT_POOL = None


def _build_round_layout(seed, n_own, n_epochs, bs, d_len, t_len):
    g = torch.Generator()
    g.manual_seed(seed)
    perms = np.concatenate([torch.randperm(n_own, generator=g).numpy() for _ in range(n_epochs)])
    return E.FastEpochLayout(perms, bs, d_len, t_len, epoch_len=n_own)


def _layout_key(org, it, n_own, n_epochs, bs):
    return (int(cfg['seed']), int(org), int(it), int(n_own), int(n_epochs), int(bs))


def _prefetch_round_layout(org, it, n_own, n_epochs, bs, d_len, t_len):
    """Build the whole-round batch layout of (organization, round) on a background thread (numpy and torch.randperm
    release the GIL): the same permutations as the foreground path — the host generator of _seeded_generators is
    seeded per (seed, organization, round) — so results do not depend on whether the prefetch was used."""
    global _LAYOUT_POOL
    key = _layout_key(org, it, n_own, n_epochs, bs)
    if key in _LAYOUTS:
        return
    if _LAYOUT_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _LAYOUT_POOL = ThreadPoolExecutor(max_workers=2)
    if len(_LAYOUTS) > 256:
        _LAYOUTS.clear()
    seed = E.he_seed(cfg['seed'], org, it, 1 << 21) & (2 ** 63 - 1)
    _LAYOUTS[key] = _LAYOUT_POOL.submit(_build_round_layout, seed, n_own, n_epochs, bs, d_len, t_len)


def _round_layout(org, it, n_own, n_epochs, bs, d_len, t_len):
    fut = _LAYOUTS.pop(_layout_key(org, it, n_own, n_epochs, bs), None)
    if fut is not None:
        return fut.result()
    return _build_round_layout(E.he_seed(cfg['seed'], org, it, 1 << 21) & (2 ** 63 - 1), n_own, n_epochs, bs, d_len, t_len)


_GENS = {}


def _seeded_generators(dev, *parts):
    """Host and device generators seeded per (experiment seed, organization, round): in device-RNG mode an
    organization's permutations, initial parameters and dropout stream do not depend on which other organizations
    this process trains, so an organization-sharded run (one process per GPU) reproduces the single-process one."""
    g = _GENS.get(dev)
    if g is None:
        g = _GENS[dev] = (torch.Generator(), torch.Generator(device=dev))
    g[0].manual_seed(E.he_seed(cfg['seed'], *parts, 1 << 21) & (2 ** 63 - 1))
    g[1].manual_seed(E.he_seed(cfg['seed'], *parts, 1 << 20) & (2 ** 63 - 1))
    return g


class Organization:
    def __init__(self, organization_id, data_split, model_name):
        self.organization_id = organization_id
        self.data_split = data_split
        self.num_items = len(data_split)
        self.model_name = model_name
        self.model_state_dict = [None for _ in range(cfg['global']['num_epochs'] + 1)]

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items() if not k.startswith('_')}
        state['model_state_dict'] = [dict(s) if s is not None else None for s in self.model_state_dict]
        return state

    # ------------------------------------------------------------------ round 0
    def initialize(self, dataset, metric, logger, iter):
        """Round-0 baseline with models.base (src/organization.py:29-138)."""
        mode = cfg['data_mode']
        implicit = cfg['target_mode'] == 'implicit'
        split = np.asarray(self.data_split, dtype=np.int64)
        bs = cfg[self.model_name[iter]]['batch_size']['train']  # the loaders use the organization's model tag
        output, target = {}, {}
        dev = _device()
        if _rng_mode() == 'reference':
            np.random.seed(cfg['seed'])  # side effect of make_data_loader (src/data.py:76)
        if 'train' in dataset:
            d = device_csr(dataset['train'].data)
            if _rng_mode() == 'reference':
                E.index_batches(1, 1, False)
                E.index_batches(1, 1, False)
            base = torch.zeros(d.shape[1], device=dev)
            count = torch.zeros(d.shape[1], device=dev)
            native.base_fit(d.indices, d.data, base, count)
            rl = d.row_len
            imp_count = float(sum(int((rl[s:s + bs] > 0).sum()) for s in range(0, d.shape[0], bs)))
            if implicit:
                count = torch.full_like(count, imp_count)
            self.model_state_dict[0] = {'base': base.cpu(), 'count': count.cpu()}
            self._base = (base, count, imp_count)
        else:
            sd = self.model_state_dict[0]
            base, count = sd['base'].to(dev), sd['count'].to(dev)
            imp_count = float(count[0]) if implicit else 0.0
        for k in dataset:
            if k == 'test' and _rng_mode() == 'reference':
                E.index_batches(1, 1, False)
            t = device_csr(dataset[k].target)
            pred = native.base_predict(base, count, t.indices, implicit, imp_count)
            rows = np.repeat(np.arange(t.shape[0]), t.row_len)
            cols = split[t.indices_host]
            pred_h = E.to_host(pred).numpy()
            output[k] = csr_matrix((pred_h, (rows, cols)), shape=_shape_target())
            target[k] = csr_matrix((np.asarray(dataset[k].target.data), (rows, cols)), shape=_shape_target())
            if k == 'train':
                self._log_initialize(dataset[k], t, pred, metric, logger, bs)
        return output, target

    def _log_initialize(self, ds, t, pred, metric, logger, bs):
        """Per-batch train metrics of the round-0 predictor, as the reference logs them (src/organization.py:46-66)."""
        mode = cfg['data_mode']
        d_len = np.diff(np.asarray(ds.data.indptr))
        tgt = torch.from_numpy(np.asarray(ds.target.data, dtype=np.float32)).to(pred.device)
        for s in range(0, t.shape[0], bs):
            e = min(t.shape[0], s + bs)
            n_in = int(d_len[s:e].sum())
            lo, hi = int(t.indptr_host[s]), int(t.indptr_host[e])
            if n_in == 0 or hi == lo:
                continue
            rows = torch.from_numpy(np.repeat(np.arange(s, e), t.row_len[s:e])).to(pred.device)
            cols = t.indices[lo:hi].long()
            inp = {'target_rating': tgt[lo:hi], 'target_' + mode: rows,
                   'target_' + ('item' if mode == 'user' else 'user'): cols}
            out = {'target_rating': pred[lo:hi]}
            out['loss'] = models.loss_fn(out['target_rating'], inp['target_rating'])
            logger.append(metric.evaluate(metric.metric_name['train'], inp, out), 'train', n_in)

    # ------------------------------------------------------------------ local training
    def _engine(self, data_m, target_m):
        if data_m.shape[0] < target_m.shape[0]:
            # cold start (src/train_recsys_assist.py:52-56): this organization's train data holds only the first rows
            # of the aligned entity while targets (and the test split) cover all of them; the engine is sized for the
            # full row range, rows beyond the data simply carry no data entries
            pad = self.__dict__.get('_padded')
            if pad is None or pad[0] is not data_m:
                ip = np.concatenate([data_m.indptr, np.full(target_m.shape[0] - data_m.shape[0], data_m.indptr[-1],
                                                            dtype=data_m.indptr.dtype)])
                pad = (data_m, csr_matrix((data_m.data, data_m.indices, ip),
                                          shape=(target_m.shape[0], data_m.shape[1])))
                self._padded = pad
            data_m = pad[1]
        d = device_csr(data_m)
        t = device_csr(target_m, with_values=False)
        bs = cfg['local']['batch_size']['train']
        enc, dec = cfg['ae']['encoder_hidden_size'], cfg['ae']['decoder_hidden_size']
        if len(enc) != 2 or len(dec) != 2 or enc[0] != dec[1] or enc[1] != dec[0]:
            raise NotImplementedError('the engine supports the reference AE shape [H1,H2]/[H2,H1] (src/utils.py:166-171)')
        # device-drawn dropout: the whole round is one plan and one graph launch when its plan buffers (~40 B per target
        # entry and epoch) stay small — 20x fewer host calls per organization, so the last organization starts early
        n_ep = int(cfg['local']['num_epochs'])
        whole = (_rng_mode() != 'reference' and 40 * t.nnz * n_ep <= (2 << 30) and t.nnz * n_ep < 2 ** 31 - 2
                 and (-(-d.shape[0] // bs) + 1) * n_ep * max(t.shape[1], d.shape[1]) < 2 ** 32)
        plan_epochs = n_ep if whole else 1
        key = (id(d), id(t), bs, plan_epochs)
        if getattr(self, '_eng_key', None) != key:
            if getattr(self, '_eng', None) is not None:
                self._eng.close()
            self._eng = E.OrgEngine(d, t, bs, enc[0], enc[1], native.LOSS_KIND[cfg['target_mode']],
                                    plan_epochs=plan_epochs)
            # all organizations of the experiment share this GPU: with many of them the decoder chunk kernel gets the
            # smaller grid that shortens the round (roundloop.DEC_BLOCKS_MANY_ORGS)
            from dmtcdr_b200 import roundloop as _rl
            n_orgs = self.__dict__.get('_orgs_on_rank') or (
                int(cfg['num_organizations']) if 'num_organizations' in cfg else 1)
            self._eng.h.set_decoder_blocks(_rl.decoder_blocks_for(n_orgs))
            self._eng.h.set_row_tile(_rl.row_tile_for(n_orgs))
            # few organizations on this GPU: a step's backward pass runs as parallel graph branches (roundloop.py)
            fan = E.os.environ.get('DMT_FANOUT')
            self._eng.h.set_fanout((n_orgs <= _rl.FANOUT_MAX_ORGS) if fan is None else fan == '1')
            self._eng_key = key
            self._residual_buf = torch.empty(t.nnz, device=_device())
        return self._eng, d, t

    def train(self, dataset, metric, logger, iter):
        """20 local Adam epochs of the AAE on the broadcast residuals (src/organization.py:140-178)."""
        if self.model_name[iter] != 'ae':
            raise TypeError("Organization.train only works with model 'ae' (as in the reference, SURVEY.md top item 1)")
        if not self._mine():
            # organization-sharded run (driver launched under torchrun): another rank trains this organization and
            # its predictions arrive through the all-gather inside Assist.update
            self.model_state_dict[iter] = None
            return
        eng, d, t = self._engine(dataset.data, dataset.target)
        n_own = dataset.data.shape[0]  # the loader walks len(dataset) rows (< the engine's row range under cold start)
        rng = _rng_mode()
        dev = _device()
        if rng == 'reference':
            np.random.seed(cfg['seed'])  # side effect of make_data_loader (src/data.py:76)
        num_users, num_items = dataset.num_users, dataset.num_items
        if rng == 'reference':
            # parameters created exactly like the reference does (same modules, same draws from torch's CPU generator)
            model = models.ae(num_users['data'], num_items['data'], num_users['target'], num_items['target'])
            flat0 = E.flat_from_state_dict(model.state_dict(), dev)
        else:
            # same initial distribution (xavier-uniform weights, zero biases, src/models/ae.py:22-28,89-96) drawn by the
            # device generator straight into the engine's layout: no host init, no upload
            from dmtcdr_b200 import roundloop
            host_gen, dev_gen = _seeded_generators(dev, self.organization_id, iter)
            flat0 = roundloop.init_flat_params(eng.n_enc, eng.n_dec, eng.H1, eng.H2, dev, dev_gen)
        res = getattr(dataset.target, '_dmt_residual_dev', None)
        if res is None:
            res = E.to_dev(np.asarray(dataset.target.data, dtype=np.float32), dev)
        self._residual_buf.copy_(res)
        eng.set_round(flat0, self._residual_buf)
        hp = dict(lr=cfg['local']['lr'], betas=tuple(cfg['local']['betas']), weight_decay=cfg['local']['weight_decay'],
                  max_norm=1.0)
        n_epochs = cfg['local']['num_epochs']
        bs = cfg['local']['batch_size']['train']
        layouts = []
        if rng == 'reference':
            losses = []
            for _ in range(n_epochs):
                lay = E.EpochLayout(E.index_batches(n_own, bs, True), eng.d_len, eng.t_len)
                keep = [torch.empty(r, eng.H2).bernoulli_(0.5) if a else torch.zeros(r, eng.H2)
                        for r, a in zip(lay.batch_rows, lay.active)]
                keep = E.to_dev(torch.cat(keep).to(torch.uint8), dev) if keep else None
                lo = torch.zeros(len(lay.active), device=dev)
                eng.enqueue_epoch(lay, keep=keep, hp=hp, loss_out=lo)
                layouts.append(lay)
                losses.append(lo)
            loss_all = torch.cat(losses)
        elif eng.plan_epochs >= n_epochs > 1:
            # the round's layout depends only on (seed, organization, round): it was prepared by the background worker
            # while the GPU ran the previous round (or is built now), and the next round's is requested right away
            lay = _round_layout(self.organization_id, iter, n_own, n_epochs, bs, eng.d_len, eng.t_len)
            if iter + 1 <= cfg['global']['num_epochs']:
                _prefetch_round_layout(self.organization_id, iter + 1, n_own, n_epochs, bs, eng.d_len, eng.t_len)
            layouts.append(lay)
            loss_all = torch.zeros(len(lay.active), device=dev)
            eng.enqueue_round(lay, E.he_seed(cfg['seed'], self.organization_id, iter, 0), hp=hp, loss_out=loss_all)
        else:
            for _ in range(n_epochs):
                layouts.append(E.FastEpochLayout(torch.randperm(n_own, generator=host_gen).numpy(), bs, eng.d_len,
                                                 eng.t_len))
            loss_all = torch.zeros(sum(len(l.active) for l in layouts), device=dev)
            seeds = [E.he_seed(cfg['seed'], self.organization_id, iter, e) for e in range(n_epochs)]
            eng.enqueue_epochs(layouts, seeds, hp=hp, loss_out=loss_all)
        flat = eng.params()
        self._eng_params_iter = iter

        def log():
            vals = E.to_host(loss_all).tolist()
            i = 0
            for lay in layouts:
                for a, n_in in zip(lay.active, lay.d_per_batch):
                    if a:
                        logger.append({metric.metric_name['train'][0]: vals[i]}, 'train', n=n_in)
                    i += 1

        sd = LazyStateDict(flat, eng, eng.n_enc, eng.n_dec, eng.H1, eng.H2, on_ready=log)
        self.model_state_dict[iter] = sd
        if 'dmt_sync' in cfg and cfg['dmt_sync']:
            sd._force()
        else:
            _PENDING.append(sd._log)  # train-loss log lines: written at the round's barrier (Assist.update)
        return

    def _mine(self):
        """False when another rank of an organization-sharded run owns this organization (set by
        Assist.make_organization from dist.assign_orgs)."""
        from dmtcdr_b200 import dist as D
        owner = self.__dict__.get('_owner_rank')
        if owner is None:
            return True
        rank, world = D.shard_context()
        return world == 1 or owner == rank

    def _remote_output(self, dataset):
        """Placeholder for an organization another rank predicts: the target's sparsity with zero values, flagged so
        that Assist.update takes this organization's vector from the exchange instead."""
        tm = dataset.target
        m = csr_matrix((np.zeros(tm.nnz, np.float32), tm.indices, tm.indptr), shape=_shape_target(), copy=False)
        m._dmt_remote = True
        return m

    def _short_output(self, out, t, n_pred):
        """Cold-start output: a CSR with the entries of the rows this organization holds (what the reference's predict
        assembles from its loader, src/organization.py:186-216); the device copy keeps the global length, NaN beyond."""
        self._eng.h.signal_current()
        cut = int(t.indptr_host[n_pred])
        ip = t.indptr_host.astype(np.int32).copy()
        ip[n_pred:] = cut
        m = csr_matrix((E.to_host(out[:cut]).numpy(), t.indices_host[:cut].astype(np.int32), ip), shape=_shape_target())
        m._dmt_pred_dev = out
        return m

    # ------------------------------------------------------------------ prediction
    def predict(self, dataset, iter):
        """Eval forward at every target position -> CSR with the sparsity of dataset.target (src/organization.py:180-217)."""
        if not self._mine():
            return self._remote_output(dataset)
        eng_data = device_csr(dataset.data)
        t = device_csr(dataset.target, with_values=False)
        if getattr(self, '_eng', None) is None:
            self._engine(dataset.data, dataset.target)
        eng = self._eng
        dev = _device()
        if _rng_mode() == 'reference':
            np.random.seed(cfg['seed'])
            nu, ni = dataset.num_users, dataset.num_items
            models.ae(nu['data'], ni['data'], nu['target'], ni['target'])  # the reference builds (and inits) a model here
            E.index_batches(1, 1, False)
        sd = self.model_state_dict[iter]
        if getattr(self, '_eng_params_iter', None) != iter:  # engine holds another round's parameters
            flat = E.flat_from_state_dict(sd, dev)
            eng.h.wait_current()
            eng.h.set_params(flat)
            self._eng_params_iter = iter
        n_pred = min(eng_data.shape[0], t.shape[0])
        short = n_pred < t.shape[0]  # cold start: rows this organization never saw stay NaN (absent in the reference)
        out = torch.full((t.nnz,), float('nan'), device=dev) if short else torch.empty(t.nnz, device=dev)
        eng.predict(eng_data, t, out)
        if short:
            return self._short_output(out, t, n_pred)
        if 'dmt_sync' in cfg and cfg['dmt_sync']:
            eng.h.signal_current()
            pred = E.to_host(out).numpy()
            if isinstance(sd, LazyStateDict):
                sd._force()
            m = csr_matrix((pred, t.indices_host.astype(np.int32, copy=False),
                            t.indptr_host.astype(np.int32, copy=False)), shape=_shape_target(), copy=False)
            m._dmt_pred_dev = out
            return m
        # device->host copy behind the organization's stream, on a side stream, into this organization's persistent
        # pinned buffer for the split; nobody waits here
        key = (t.nnz, id(t))
        pins = self.__dict__.setdefault('_pred_pins', {})
        slot = pins.get(key)
        if slot is None:
            slot = pins[key] = {'pin': torch.empty(t.nnz, dtype=torch.float32, pin_memory=True), 'last': None}
        prev = slot['last']() if slot['last'] is not None else None
        if prev is not None:
            prev.data  # an earlier, still unread output shares the pinned buffer: give it its values first
        eng.h.signal_current()  # Assist.update reads `out` on the current stream: order it behind this organization
        d2h = _d2h_stream(dev)
        with torch.cuda.stream(d2h):
            eng.h.signal_current()  # the side stream waits for this organization's stream
            slot['pin'].copy_(out, non_blocking=True)
            done = d2h.record_event()
        out.record_stream(d2h)
        E.XFER["d2h"] += out.numel() * out.element_size()
        host = np.empty(t.nnz, dtype=np.float32)
        m = LazyCSR((host, t.indices_host.astype(np.int32, copy=False), t.indptr_host.astype(np.int32, copy=False)),
                    shape=_shape_target(), copy=False)
        pin = slot['pin']
        dest = m.__dict__['_dmt_data']  # the array scipy actually kept

        def force():
            done.synchronize()
            np.copyto(dest, pin.numpy())

        m.__dict__['_dmt_force'] = force
        slot['last'] = weakref.ref(m)
        m._dmt_pred_dev = out
        return m
