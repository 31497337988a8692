"""Standalone experiment loop over the drop-in API (what reference src/train_recsys_assist.py:40-93 does), used by
the GPU parity tests, smoke() and bench.py's end-to-end leg on boxes where the reference's drivers do not exist.

It builds dataset objects with the attribute surface the hot path reads (``.data`` / ``.target`` scipy CSR,
``.num_users`` / ``.num_items``, ``.transform.transforms[0]``, ``.item_attr`` / ``.user_profile``), splits the
columns into organizations like src/data.py:200-274, and drives ``Assist`` / ``Organization`` exactly in the
reference's order so that, in ``dmt_rng='reference'`` mode, torch's generator is consumed identically.
"""
import copy
import types

import numpy as np
import torch

from . import use_dropin
from .config import cfg, make_cfg
from .metrics import Logger, Metric


class SplitDataset:
    """One split ('train' / 'test') in the orientation of cfg['data_mode'] (rows = aligned entity)."""

    def __init__(self, data, target, data_mode, item_attr=None, user_profile=None):
        self.data, self.target, self.data_mode = data, target, data_mode
        if item_attr is not None:
            self.item_attr = {'data': item_attr, 'target': item_attr}
        if user_profile is not None:
            self.user_profile = {'data': user_profile, 'target': user_profile}
        sizes = types.SimpleNamespace(num_users=self.num_users, num_items=self.num_items)
        self.transform = types.SimpleNamespace(transforms=[sizes])

    def __len__(self):
        return self.data.shape[0]

    def _sizes(self, axis_rows):
        a = 0 if axis_rows else 1
        return {'data': self.data.shape[a], 'target': self.target.shape[a]}

    @property
    def num_users(self):
        return self._sizes(self.data_mode == 'user')

    @property
    def num_items(self):
        return self._sizes(self.data_mode == 'item')


def fetch_dataset(data):
    """{'train','test'} like the reference's datasets: test 'data' is the train matrix (src/datasets/movielens.py:367-371)."""
    (trd, trt), (ted, tet) = data.split(cfg['target_mode'])
    mats = {'train': (trd, trt), 'test': (ted, tet)}
    if cfg['data_mode'] == 'item':
        mats = {k: (a.T.tocsr(), b.T.tocsr()) for k, (a, b) in mats.items()}
    out = {}
    for k, (a, b) in mats.items():
        a.sort_indices()
        b.sort_indices()
        out[k] = SplitDataset(a, b, cfg['data_mode'], data.item_attr, data.user_profile)
    return out


def process_dataset(dataset):
    cfg['data_size'] = {'train': len(dataset['train']), 'test': len(dataset['test'])}
    cfg['num_users'], cfg['num_items'] = dataset['train'].num_users, dataset['train'].num_items
    cfg['info_size'] = None  # side information (info=1) is not wired into the engine path


def split_dataset(dataset):
    """Column split into organizations (src/data.py:200-242): 'genre' = one multinomial draw per column over its
    genre indicator, redrawn until every organization is non-empty in all four matrices; 'random-K' = randperm chunks."""
    K = cfg['num_organizations']
    mode = cfg['data_split_mode']
    if 'genre' in mode:
        if cfg['data_mode'] != 'user':
            raise NotImplementedError
        attr = torch.tensor(dataset['train'].item_attr['data'])
        attr[attr.sum(-1) == 0] = 1
        while True:
            idx = torch.multinomial(attr, 1).view(-1).numpy()
            split = [np.where(idx == i)[0] for i in range(K)]
            mats = [dataset[k].data for k in ('train', 'test')] + [dataset[k].target for k in ('train', 'test')]
            if all(len(s) > 0 and all(m[:, s].nnz > 0 for m in mats) for s in split):
                return [torch.tensor(s) for s in split]
    if 'random' in mode:
        n = dataset['train'].data.shape[1]
        chunks = list(torch.randperm(n).split(n // K))
        return chunks[:K - 1] + [torch.cat(chunks[K - 1:])]
    raise ValueError('Not valid data split mode')


def make_split_dataset(dataset, data_split):
    out = []
    for s in data_split:
        cols = s.numpy()
        out.append({k: SplitDataset(dataset[k].data[:, cols].tocsr(), dataset[k].target[:, cols].tocsr(),
                                    cfg['data_mode']) for k in dataset})
    return out


def pair_batch(ds, rows):
    """COO batch of the listed rows in the reference's PairInput format (src/data.py:84-137): the aligned id repeated
    per entry, concatenated in the given row order; train keys and target_* keys."""
    mode = cfg['data_mode']
    other = 'item' if mode == 'user' else 'user'
    rows = np.asarray(rows, dtype=np.int64)
    out = {}
    for pre, m in (('', ds.data), ('target_', ds.target)):
        starts, ends = m.indptr[rows], m.indptr[rows + 1]
        cnt = (ends - starts).astype(np.int64)
        pos = np.repeat(starts.astype(np.int64) - np.concatenate([[0], np.cumsum(cnt)[:-1]]), cnt) + np.arange(cnt.sum())
        out[pre + mode] = torch.from_numpy(np.repeat(rows, cnt))
        out[pre + other] = torch.from_numpy(m.indices[pos].astype(np.int64))
        out[pre + 'rating'] = torch.from_numpy(m.data[pos].astype(np.float32))
    return out


def initialize(dataset, assist, organization, metric, logger):
    """Round 0: every organization's base predictor, assembled into the global output / target matrices
    (src/train_recsys_assist.py:98-141)."""
    from scipy.sparse import csr_matrix

    parts = {k: {'o': [], 't': []} for k in dataset[0]}
    for i in range(len(dataset)):
        out_i, tgt_i = organization[i].initialize(dataset[i], metric, logger, 0)
        for k in dataset[0]:
            parts[k]['o'].append(out_i[k].tocoo())
            parts[k]['t'].append(tgt_i[k].tocoo())
    shape = (cfg['num_users']['target'], cfg['num_items']['target']) if cfg['data_mode'] == 'user' else \
        (cfg['num_items']['target'], cfg['num_users']['target'])
    for k in dataset[0]:
        for name, store in (('o', assist.organization_output), ('t', assist.organization_target)):
            coo = parts[k][name]
            store[0][k] = csr_matrix((np.concatenate([c.data for c in coo]),
                                      (np.concatenate([c.row for c in coo]), np.concatenate([c.col for c in coo]))),
                                     shape=shape)
    logger.safe(False)
    logger.reset()


def evaluate(assist, metric, logger, epoch):
    """Global test metrics in row blocks of the test batch size (src/train_recsys_assist.py:175-217)."""
    import models

    F = assist.organization_output[epoch]['test']
    y = assist.organization_target[0]['test']
    if 'cs' in cfg:  # cold-start runs are scored on organization 0's columns (src/train_recsys_assist.py:180-182)
        cols0 = np.asarray(assist.data_split[0])
        F, y = F[:, cols0], y[:, cols0]
    bs = cfg[cfg['model_name']]['batch_size']['test']
    mode = cfg['data_mode']
    for s in range(0, F.shape[0], bs):
        lo, hi = y.indptr[s], y.indptr[min(F.shape[0], s + bs)]
        if hi == lo:
            continue
        rows = torch.from_numpy(np.repeat(np.arange(s, min(F.shape[0], s + bs)),
                                          np.diff(y.indptr[s:min(F.shape[0], s + bs) + 1])).astype(np.int64))
        cols = torch.from_numpy(y.indices[lo:hi].astype(np.int64))
        out = {'target_rating': torch.from_numpy(F.data[lo:hi])}
        inp = {'target_rating': torch.from_numpy(y.data[lo:hi]), 'target_' + mode: rows,
               'target_' + ('item' if mode == 'user' else 'user'): cols}
        out['loss'] = models.loss_fn(out['target_rating'], inp['target_rating'])
        logger.append(metric.evaluate(metric.metric_name['test'], inp, out), 'test', n=int(hi - lo))
    return {k: float(v) for k, v in logger.mean.items() if k.startswith('test/')}


def joint_test_device(local_model, local_dataset, data_split, y_test, topk=10):
    """The joint / alone drivers' test() loop (src/train_recsys_joint.py:153-199, src/train_recsys_alone.py:164-203) on
    the device: every organization's local model (after ``models.distribute``) predicts its own block of the test
    targets — ONE forward over all of the organization's test entries instead of one per 'test' batch — the predictions
    are scattered into the global test CSR's storage order, and one `dmt_eval_blocks` launch produces the per-block
    Loss / RMSE / NDCG@k sums with the reference's weighting (per-organization losses weighted by entry count ==
    the global mean; RMSE / NDCG per block of `batch_size['test']` aligned rows over the organizations' concatenated
    entries, weighted by the block's entry count). Returns {'test/Loss', 'test/RMSE' | 'test/NDCG'}.

    local_dataset[i]['test'] is organization i's split (``.target`` = its columns of y_test); ae models take the
    engine path instead (Organization.predict)."""
    from . import engine as E
    from . import native

    dev = cfg['device']
    mode = cfg['data_mode']
    y = y_test.tocsr()
    y.sort_indices()
    n_cols = y.shape[1]
    owner = np.full(n_cols, -1, np.int64)
    local = np.zeros(n_cols, np.int64)
    for i, sp in enumerate(data_split):
        sp = np.asarray(sp, dtype=np.int64)
        owner[sp] = i
        local[sp] = np.arange(len(sp))
    pred = torch.empty(y.nnz, device=dev)
    rows_all = np.repeat(np.arange(y.shape[0], dtype=np.int64), np.diff(y.indptr))
    with torch.no_grad():
        for i, model in enumerate(local_model):
            pos = np.flatnonzero(owner[y.indices] == i)
            if pos.size == 0:
                continue
            model.train(False)
            batch = {'target_' + mode: torch.from_numpy(rows_all[pos]).to(dev),
                     'target_' + ('item' if mode == 'user' else 'user'): torch.from_numpy(local[y.indices[pos]]).to(dev),
                     'target_rating': torch.from_numpy(y.data[pos].astype(np.float32)).to(dev)}
            out = model(batch)
            pred[torch.from_numpy(pos).to(dev)] = out['target_rating']
    state = E.MtalState({'test': y}, [np.asarray(sp, dtype=np.int64) for sp in data_split], cfg['target_mode'], dev)
    bs = cfg[cfg['model_name']]['batch_size']['test']
    out = state.evaluate(pred, 'test', bs, topk)
    if cfg['target_mode'] == 'implicit':
        # The driver concatenates the organizations' batches with their LOCAL ids of the split entity and hands them to
        # NDCG (src/train_recsys_joint.py:186-196), whose dense scatter `output_[user_idx, item_idx] = output`
        # (src/metrics/metrics.py:66-76) keeps ONE entry per (aligned id, local id): on the CPU the last writer, i.e.
        # the organization with the highest index. Same collapse here; block weights stay the uncollapsed entry counts
        # (logger.append(evaluation, 'test', input_size)), the per-block mean runs over the aligned ids present.
        n_local = max(len(sp) for sp in data_split)
        key = rows_all * n_local + local[y.indices]
        order = np.lexsort((owner[y.indices], key))           # by key, then organization
        last = np.r_[key[order][1:] != key[order][:-1], True]  # highest organization of every key
        win = order[last]                                      # ascending key == row-major, ascending local id
        cnt = np.bincount(rows_all[win], minlength=y.shape[0])
        ip = np.concatenate([[0], np.cumsum(cnt)])
        edges = np.arange(0, y.shape[0] + bs, bs).clip(max=y.shape[0])
        m_full = (y.indptr[edges[1:]] - y.indptr[edges[:-1]]).astype(np.float64)
        rows_nz = np.add.reduceat(np.diff(ip) > 0, edges[:-1]).astype(np.float64)
        loc_win = local[y.indices[win]]
        k = np.array([min(topk, len(np.unique(loc_win[ip[a]:ip[b]]))) for a, b in zip(edges[:-1], edges[1:])], np.int32)
        win_d = torch.from_numpy(win).to(dev)
        tgt = torch.from_numpy(y.data.astype(np.float32)).to(dev)
        sums = E.to_host(native.eval_blocks(E.to_dev(ip.astype(np.int32), dev), pred[win_d].contiguous(),
                                            tgt[win_d].contiguous(), y.shape[0], bs, state.loss_kind,
                                            E.to_dev(k, dev))).double().numpy()
        ok = m_full > 0
        w = m_full[ok] / m_full[ok].sum()
        out['test/NDCG'] = float((sums[ok, 2] / rows_nz[ok] * w).sum())
    return out


def run_assist_experiment(data, control_name, seed=0, local_epochs=None, rounds=None, rng='device', keep_objects=False,
                          on_round=None, materialize_state_dicts=False):
    """Whole MTAL experiment through the drop-in API. Returns per-round global outputs and test metrics.
    materialize_state_dicts: read every organization's state_dict at the end of every round, as the reference's
    per-round checkpoint of whole objects does (src/train_recsys_assist.py:87-89; src/organization.py:177 copies them
    to the CPU in train()); by default they stay on the device until somebody reads them."""
    make_cfg(control_name, device='cuda', seed=seed)
    cfg['dmt_rng'] = rng
    if local_epochs is not None:
        cfg['local']['num_epochs'] = local_epochs
    if rounds is not None:
        cfg['global']['num_epochs'] = rounds
    models, organization_mod, assist_mod = use_dropin()
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    dataset = fetch_dataset(data)
    process_dataset(dataset)
    data_split = split_dataset(dataset)
    dataset = make_split_dataset(dataset, data_split)
    if 'cs' in cfg:  # cold start: organization 0 keeps the first int(n*cs) aligned rows (src/train_recsys_assist.py:52-56)
        start_size = int(len(dataset[0]['train']) * cfg['cs'])
        dataset[0]['train'].data = dataset[0]['train'].data[:start_size]
        dataset[0]['train'].target = dataset[0]['train'].target[:start_size]
    assist = assist_mod.Assist(data_split)
    organization = assist.make_organization()
    names = ['Loss', 'RMSE'] if cfg['target_mode'] == 'explicit' else ['Loss', 'NDCG']
    metric = Metric({'train': names, 'test': names})
    logger = Logger()
    initialize(dataset, assist, organization, metric, logger)
    metrics = {0: evaluate(assist, metric, logger, 0)}
    logger.reset()
    import os, time
    trace = os.environ.get('DMT_TRACE') == '1'
    for t in range(1, cfg['global']['num_epochs'] + 1):
        tm = [time.perf_counter()]
        dataset = assist.make_dataset(dataset, t)
        tm.append(time.perf_counter())
        for i in range(len(organization)):
            organization[i].train(dataset[i]['train'], metric, logger, t)
        tm.append(time.perf_counter())
        outs = [{k: organization[i].predict(dataset[i][k], t) for k in dataset[i]} for i in range(len(dataset))]
        tm.append(time.perf_counter())
        assist.update(outs, t)
        if materialize_state_dicts:
            for o in organization:
                sd = o.model_state_dict[t]
                if sd is not None:
                    len(sd)
        tm.append(time.perf_counter())
        metrics[t] = evaluate(assist, metric, logger, t)
        tm.append(time.perf_counter())
        logger.reset()
        if trace:
            names = ['make_dataset', 'train', 'predict', 'update', 'evaluate']
            print('round', t, {n: round(1e3 * (b - a), 1) for n, a, b in zip(names, tm[:-1], tm[1:])}, flush=True)
        if on_round is not None:
            on_round(t)
    res = {'F': [{k: m[k].data.copy() for k in m} for m in assist.organization_output], 'metrics': metrics,
           'data_split': [s.numpy() for s in data_split], 'y': assist.organization_target[0],
           'ar_state_dict': assist.ar_state_dict}
    if keep_objects:
        res.update(assist=assist, organization=organization, dataset=dataset)
    return res
