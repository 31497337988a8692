"""Drop-in ``assist`` module: the MTAL coordinator ``Assist`` with the reference's surface
(src/assist.py:13-179: ``make_organization`` / ``make_dataset`` / ``update`` / ``reset`` and the attributes
``data_split``, ``num_organizations``, ``model_name``, ``ar_state_dict``, ``organization_output``,
``organization_target``). Host objects (scipy CSR, CPU state dicts) at the API edge, device arithmetic inside.
"""
import numpy as np
import torch
from scipy.sparse import csr_matrix

from dmtcdr_b200 import engine as E
from dmtcdr_b200 import native
from dmtcdr_b200.config import cfg
from dmtcdr_b200.privacy import make_privacy

from organization import Organization, _device


class Assist:
    def __init__(self, data_split):
        self.data_split = data_split
        self.num_organizations = len(data_split)
        self.model_name = self.make_model_name()
        self.ar_state_dict = [[None for _ in range(cfg['num_organizations'])] for _ in
                              range(cfg['global']['num_epochs'] + 1)]
        self.reset()

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith('_')}

    def reset(self):
        n = cfg['global']['num_epochs'] + 1
        self.organization_output = [{k: None for k in cfg['data_size']} for _ in range(n)]
        self.organization_target = [{k: None for k in cfg['data_size']} for _ in range(n)]
        self._state = None
        self._F_dev = {}

    def make_model_name(self):
        return [[cfg['model_name'] for _ in range(cfg['global']['num_epochs'] + 1)]
                for _ in range(self.num_organizations)]

    def make_organization(self):
        orgs = [Organization(i, self.data_split[i], self.model_name[i]) for i in range(self.num_organizations)]
        sh = self._sharding()
        if sh is not None:
            for r, mine in enumerate(sh['mine']):
                for i in mine:
                    orgs[i]._owner_rank = r
                    orgs[i]._orgs_on_rank = len(mine)
        return orgs

    def _sharding(self):
        """Organization -> rank map when the driver runs one process per GPU (torch.distributed initialised, e.g. the
        unmodified train_recsys_assist.py under torchrun): the reference's loops over organizations
        (src/train_recsys_assist.py:148-149,169-171) then do real work only for this rank's organizations and
        Assist.update all-gathers the prediction vectors over NCCL (dist.py). None in a single-process run."""
        from dmtcdr_b200 import dist as D
        rank, world = D.shard_context()
        if world == 1:
            return None
        sh = self.__dict__.get('_shard')
        if sh is None or sh['world'] != world:
            mine, chunk, org_row = D.assign_orgs([len(s) for s in self.data_split], world)
            sh = self._shard = {'rank': rank, 'world': world, 'mine': mine, 'chunk': chunk, 'org_row': org_row}
        return sh

    # ------------------------------------------------------------------ device state
    def _mtal(self):
        if self.__dict__.get('_state') is None:
            y = {k: self.organization_target[0][k] for k in self.organization_target[0]
                 if self.organization_target[0][k] is not None}
            for m in y.values():
                m.sort_indices()
            split = [np.asarray(s.cpu() if isinstance(s, torch.Tensor) else s, dtype=np.int64) for s in self.data_split]
            sh = self._sharding()
            self._state = E.MtalState(y, split, cfg['target_mode'], _device(),
                                      o_rows=sh['chunk'] * sh['world'] if sh else None,
                                      org_row=sh['org_row'] if sh else None)
            self._F_dev = {}
        return self._state

    def _F(self, iter, split):
        """Global prediction of a round on the device (uploaded once from the host CSR if we did not produce it)."""
        key = (iter, split)
        if key not in self._F_dev:
            m = self.organization_output[iter][split]
            st = self._mtal()
            ref = st.y[split]
            if m.nnz != ref.nnz or not np.array_equal(m.indptr, ref.indptr_host):
                raise ValueError('organization_output and organization_target must share one sparsity pattern')
            self._F_dev[key] = E.to_dev(np.asarray(m.data, dtype=np.float32), st.device)
        return self._F_dev[key]

    # ------------------------------------------------------------------ make_dataset
    def make_dataset(self, dataset, iter):
        """Broadcast the pseudo-residuals r = -dL/dF of round iter-1 as every organization's new target
        (src/assist.py:43-79)."""
        st = self._mtal()
        clamp = cfg['data_name'] in ['Douban', 'Amazon'] and not (
            cfg['data_name'] == 'Douban' and cfg['data_mode'] == 'item' and cfg['target_mode'] == 'explicit')
        for k in dataset[0]:
            res_dev = st.residual(self._F(iter - 1, k), k, clamp)
            res = E.to_host(res_dev).numpy()
            if 'pl' in cfg and cfg['pl'] != 'none':
                res = make_privacy(res, cfg['pl_mode'], cfg['pl_param'])
                res_dev = E.to_dev(res, st.device)
            ref = st.y[k]
            if cfg['data_mode'] == 'user':
                shape = (cfg['num_users']['target'], cfg['num_items']['target'])
            elif cfg['data_mode'] == 'item':
                shape = (cfg['num_items']['target'], cfg['num_users']['target'])
            else:
                raise ValueError('Not valid data mode')
            for i in range(len(dataset)):
                ds = dataset[i][k]
                if hasattr(ds, 'user_profile') and 'target' in ds.user_profile:
                    del ds.user_profile['target']
                if hasattr(ds, 'item_attr') and 'target' in ds.item_attr:
                    del ds.item_attr['target']
                # every organization gets its OWN csr object (as in the reference) over shared index arrays
                tgt = csr_matrix((res, ref.indices_host.astype(np.int32, copy=False),
                                  ref.indptr_host.astype(np.int32, copy=False)), shape=shape, copy=False)
                tgt._dmt_residual_dev = res_dev
                ds.target = tgt
                tr = getattr(getattr(ds, 'transform', None), 'transforms', None)
                if tr:
                    if cfg['data_mode'] == 'user':
                        tr[0].num_items['target'] = cfg['num_items']['target']
                    else:
                        tr[0].num_users['target'] = cfg['num_users']['target']
        return dataset

    # ------------------------------------------------------------------ update
    def update(self, organization_outputs, iter):
        """F_t = F_{t-1} + eta[idx] * sum_j softmax(w)_j out_j for every owner, with the optional L-BFGS fit of eta / w
        on the train split and partial alignment (src/assist.py:81-179)."""
        # cold start (12-field control names): organization 0 predicted only the aligned rows it holds; the rest of its
        # vector is NaN, which selects the softmax(w[1:]) branch of the combine (src/assist.py:109-117,150-157)
        cold = 'cs' in cfg and float(cfg['cs']) < 1
        st = self._mtal()
        import organization as _org_mod
        for k in organization_outputs[0]:
            for j, out in enumerate(organization_outputs):
                m = out[k]
                if getattr(m, '_dmt_remote', False):
                    continue  # another rank's organization: its row arrives by the all-gather below
                dev_vals = getattr(m, '_dmt_pred_dev', None)
                if dev_vals is None:
                    vals = np.asarray(m.data, dtype=np.float32)
                    if m.nnz != st.y[k].nnz:
                        if not (cold and j == 0 and m.nnz < st.y[k].nnz
                                and np.array_equal(m.indices, st.y[k].indices_host[:m.nnz])):
                            raise ValueError('organization output {} does not match the target sparsity'.format(j))
                        vals = np.concatenate([vals, np.full(st.y[k].nnz - m.nnz, np.nan, np.float32)])
                    dev_vals = E.to_dev(vals, st.device)
                elif dev_vals.numel() != st.y[k].nnz:
                    raise ValueError('organization output {} does not match the target sparsity'.format(j))
                st.o_row(k, j).copy_(dev_vals)
        sh = self._sharding()
        if sh is not None:
            from dmtcdr_b200 import dist as D
            D.exchange_outputs(st.O_full, sh['chunk'], sh['rank'], sh['world'])
        # the round's barrier: everything the organizations deferred (train-loss log lines) is written now, before
        # the driver evaluates and resets its logger
        _org_mod.flush_pending()
        a = cfg['assist']
        match_rate = a['match_rate'] if 'match_rate' in a else 1.0
        F_prev = {k: self._F(iter - 1, k) for k in organization_outputs[0]}
        F_next, fitted = st.update(F_prev, a['ar'], a['ar_mode'], a['aw_mode'], match_rate, cold=cold)
        for i, (rate, weight) in enumerate(fitted):
            self.ar_state_dict[iter][i] = {'assist_rate': rate.cpu(), 'assist_weight': weight.cpu()}
        if cfg['data_mode'] == 'user':
            shape = (cfg['num_users']['target'], cfg['num_items']['target'])
        elif cfg['data_mode'] == 'item':
            shape = (cfg['num_items']['target'], cfg['num_users']['target'])
        else:
            raise ValueError('Not valid data mode')
        pins = self.__dict__.setdefault('_F_pins', {})
        for k, F in F_next.items():
            ref = st.y[k]
            self._F_dev[(iter, k)] = F
            if 'dmt_sync' in cfg and cfg['dmt_sync']:
                self.organization_output[iter][k] = csr_matrix(
                    (E.to_host(F).numpy(), ref.indices_host.astype(np.int32, copy=False),
                     ref.indptr_host.astype(np.int32, copy=False)), shape=shape, copy=False)
            else:
                # the host copy of F_t travels on the side stream and is materialised on first read: the driver reads
                # the test split every round (its test() loop), the train split only through make_dataset, which takes
                # the device copy above
                self.organization_output[iter][k] = _org_mod.lazy_csr(F, ref.indices_host, ref.indptr_host, shape,
                                                                      pins.setdefault(k, {}))
        return
