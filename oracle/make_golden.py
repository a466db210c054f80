#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE MODEL CLASSES.  (test infrastructure)

Runs only in the build container: it needs ``/root/reference`` (read-only mount) on ``sys.path`` next to
``oracle/recbole_shim`` (stub of the un-vendored ``recbole==1.0.1``).  The GPU box has neither, which is why
the outputs are committed as small fixtures.  Usage::

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

For every case we store: the scalar hyper-parameters, every model parameter (state_dict, so the checker
does not depend on torch's RNG stream), the batch, and the reference's outputs -- loss, dense gradients of
every parameter after ``loss.backward()``, and ``predict`` scores.

Harness-side compatibility patches (none of them touches reference arithmetic):
  * ``np.NINF = -np.inf``   -- removed in NumPy 2; DTCDR fills dead rows with it before xavier init overwrites them.
  * ``BiTGCF.get_norm_adj_mat`` is replaced by ``oracle.cdr_oracle.bitgcf_norm_adj`` because the reference uses
    the private ``dok_matrix._update`` (bitgcf.py:101) which SciPy >= 1.13 no longer has.  The replacement is
    cross-checked below against a slow dict-based build that follows bitgcf.py:96-110 step by step.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('XDR_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'recbole_shim'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np
import scipy.sparse as sp
import torch

np.NINF = -np.inf  # NumPy-2 compatibility for dtcdr.py:55-59

from recbole_cdr.model.cross_domain_recommender.emcdr import EMCDR  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.cmf import CMF  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.conet import CoNet  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.dtcdr import DTCDR  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.bitgcf import BiTGCF  # noqa: E402
from oracle import cdr_oracle  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


class _Domain:
    def __init__(self, prefix, n_users, n_items):
        self.uid_field = f'{prefix}_user_id'
        self.iid_field = f'{prefix}_item_id'
        self.label_field = f'{prefix}_label'
        self._num = {self.uid_field: n_users, self.iid_field: n_items}

    def num(self, field):
        return self._num[field]


class FakeDataset:
    """The attributes CrossDomainRecommender.__init__ reads (crossdomain_recommender.py:24-45) on the joint id
    layout of data/dataset.py:344-445: 0=[PAD], [1,n_ov) overlapped, then target-only, then source-only."""

    def __init__(self, n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, seed=0, edges_per_user=4):
        self.num_overlap_user, self.num_overlap_item = n_ov_u, n_ov_i
        self.num_target_only_user, self.num_source_only_user = n_tgt_u, n_src_u
        self.num_target_only_item, self.num_source_only_item = n_tgt_i, n_src_i
        self.num_total_user = n_ov_u + n_tgt_u + n_src_u
        self.num_total_item = n_ov_i + n_tgt_i + n_src_i
        self.source_domain_dataset = _Domain('source', n_ov_u + n_src_u, n_ov_i + n_src_i)
        self.target_domain_dataset = _Domain('target', n_ov_u + n_tgt_u, n_ov_i + n_tgt_i)
        self.overlap_id_field = 'overlap'
        rng = np.random.RandomState(seed)
        self.edges = {}
        for dom in ('source', 'target'):
            users, items = self.valid_ids(dom)
            r = np.repeat(users, edges_per_user)
            c = rng.choice(items, size=r.shape[0])
            e = np.unique(np.stack([r, c], 1), axis=0)
            self.edges[dom] = (e[:, 0], e[:, 1])

    def valid_ids(self, domain):
        ou, oi = self.num_overlap_user, self.num_overlap_item
        tu, ti = ou + self.num_target_only_user, oi + self.num_target_only_item
        if domain == 'source':
            return (np.concatenate([np.arange(1, ou), np.arange(tu, self.num_total_user)]),
                    np.concatenate([np.arange(1, oi), np.arange(ti, self.num_total_item)]))
        return np.arange(1, tu), np.arange(1, ti)

    def inter_matrix(self, form='coo', value_field=None, domain='source'):
        r, c = self.edges[domain]
        m = sp.coo_matrix((np.ones(len(r)), (r, c)), shape=(self.num_total_user, self.num_total_item))
        return m.tocsr() if form == 'csr' else m


def base_config(**kw):
    cfg = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': 'cpu'}
    cfg.update(kw)
    return cfg


def make_batch(ds, domain, B, rng, pairwise=False):
    users, items = ds.valid_ids(domain)
    b = {f'{domain}_user_id': torch.from_numpy(rng.choice(users, B)).long(),
         f'{domain}_item_id': torch.from_numpy(rng.choice(items, B)).long()}
    if pairwise:
        b[f'neg_{domain}_item_id'] = torch.from_numpy(rng.choice(items, B)).long()
    else:
        b[f'{domain}_label'] = torch.from_numpy((rng.rand(B) < 0.5).astype(np.float32))
    return b


def run_and_pack(model, batch, extra=None, predict=True):
    model.zero_grad()
    loss = model.calculate_loss(batch)
    out = {}
    if isinstance(loss, tuple):
        for k, l in enumerate(loss):
            out[f'loss{k}'] = l.detach().numpy().astype(np.float32).reshape(-1)
        total = sum(loss)
    else:
        out['loss0'] = loss.detach().numpy().astype(np.float32).reshape(-1)
        total = loss
    total.sum().backward()
    for name, p in model.named_parameters():
        out['param/' + name] = p.detach().numpy().copy()
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out['grad/' + name] = g.detach().numpy().copy()
    for k, v in batch.items():
        out['batch/' + k] = v.numpy()
    if predict:
        with torch.no_grad():
            out['predict'] = model.predict(batch).detach().numpy()
    if extra:
        out.update(extra)
    return out


def save(name, d):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **d)
    print(f'{name:28s} loss={[float(d[k][0]) for k in sorted(d) if k.startswith("loss")]}  '
          f'{os.path.getsize(path) / 1024:.0f} KiB')


def sizes_scalar(ds):
    return {'meta/n_ov_u': ds.num_overlap_user, 'meta/n_tgt_u': ds.num_target_only_user,
            'meta/n_src_u': ds.num_source_only_user, 'meta/n_ov_i': ds.num_overlap_item,
            'meta/n_tgt_i': ds.num_target_only_item, 'meta/n_src_i': ds.num_source_only_item}


def slow_norm_adj(inter, n_users, n_items):
    """bitgcf.py:96-116 followed literally with a python dict + COO assignment (no dok_matrix._update)."""
    n = n_users + n_items
    t = inter.transpose()
    data = dict(zip(zip(inter.row, inter.col + n_users), [1] * inter.nnz))
    data.update(dict(zip(zip(t.row + n_users, t.col), [1] * t.nnz)))
    keys = np.array(list(data.keys()))
    A = sp.coo_matrix((np.ones(len(keys), dtype=np.float32), (keys[:, 0], keys[:, 1])), shape=(n, n)).tocsr()
    sumArr = (A > 0).sum(axis=1)
    diag = np.array(sumArr.flatten())[0] + 1e-7
    diag = np.power(diag, -0.5)
    D = sp.diags(diag)
    L = sp.coo_matrix(D * A * D)
    return torch.sparse_coo_tensor(torch.LongTensor(np.array([L.row, L.col])), torch.FloatTensor(L.data),
                                   torch.Size(L.shape)).coalesce()


def main():
    D = 64
    # ---------------- EMCDR: user-overlap layout (items disjoint: num_overlap_item == 1) ----------------
    ds_u = FakeDataset(41, 30, 35, 1, 50, 60)
    for lfm, pairwise in (('BPR', True), ('MF', False)):
        for phase, dom in (('SOURCE', 'source'), ('TARGET', 'target')):
            torch.manual_seed(2022)
            cfg = base_config(latent_factor_model=lfm, source_embedding_size=D, target_embedding_size=D,
                              reg_weight=0.01, mapping_function='non_linear', mlp_hidden_size=[128])
            m = EMCDR(cfg, ds_u)
            m.set_phase(phase)
            rng = np.random.RandomState(7)
            batch = make_batch(ds_u, dom, 96, rng, pairwise=pairwise)
            extra = sizes_scalar(ds_u)
            extra['meta/reg_weight'] = 0.01
            save(f'emcdr_{lfm.lower()}_{phase.lower()}', run_and_pack(m, batch, extra))
    # map phase (non_linear and linear), idx keeps the reference's [b, 1] shape and includes PAD 0
    for mf in ('non_linear', 'linear'):
        torch.manual_seed(2022)
        cfg = base_config(latent_factor_model='BPR', source_embedding_size=D, target_embedding_size=D,
                          reg_weight=0.01, mapping_function=mf, mlp_hidden_size=[128])
        m = EMCDR(cfg, ds_u)
        m.set_phase('OVERLAP')
        rng = np.random.RandomState(11)
        idx = torch.from_numpy(rng.permutation(ds_u.num_overlap_user)[:32].reshape(-1, 1)).long()
        batch = {'overlap': idx}
        pb = make_batch(ds_u, 'target', 80, rng)  # predict batch mixes overlapped and target-only users
        with torch.no_grad():
            pred = m.predict(pb).numpy()
        extra = sizes_scalar(ds_u)
        extra.update({'pbatch/' + k: v.numpy() for k, v in pb.items()})
        extra['predict_overlap_phase'] = pred
        save(f'emcdr_map_{mf}', run_and_pack(m, batch, extra, predict=False))
    # item-overlap layout: predict path of emcdr.py:200-205 and map loss on items
    ds_i = FakeDataset(1, 40, 45, 31, 30, 33)
    torch.manual_seed(2022)
    cfg = base_config(latent_factor_model='BPR', source_embedding_size=D, target_embedding_size=D,
                      reg_weight=0.01, mapping_function='non_linear', mlp_hidden_size=[128])
    m = EMCDR(cfg, ds_i)
    m.set_phase('OVERLAP')
    rng = np.random.RandomState(13)
    idx = torch.from_numpy(rng.permutation(ds_i.num_overlap_item)[:24].reshape(-1, 1)).long()
    pb = make_batch(ds_i, 'target', 80, rng)
    with torch.no_grad():
        pred = m.predict(pb).numpy()
    extra = sizes_scalar(ds_i)
    extra.update({'pbatch/' + k: v.numpy() for k, v in pb.items()})
    extra['predict_overlap_phase'] = pred
    save('emcdr_map_items', run_and_pack(m, {'overlap': idx}, extra, predict=False))

    # ---------------- CMF: both users and items may overlap ----------------
    ds_b = FakeDataset(25, 30, 28, 21, 30, 26)
    torch.manual_seed(2022)
    m = CMF(base_config(embedding_size=D, alpha=0.3, gamma=0.02, **{'lambda': 0.05}), ds_b)
    rng = np.random.RandomState(17)
    batch = make_batch(ds_b, 'source', 96, rng)
    batch.update(make_batch(ds_b, 'target', 80, rng))
    extra = sizes_scalar(ds_b)
    extra.update({'meta/alpha': 0.3, 'meta/lambda': 0.05, 'meta/gamma': 0.02})
    save('cmf_both', run_and_pack(m, batch, extra))

    # ---------------- CoNet (user overlap and item overlap) ----------------
    for tag, ds in (('users', ds_u), ('items', ds_i)):
        torch.manual_seed(2022)
        m = CoNet(base_config(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8]), ds)
        rng = np.random.RandomState(19)
        batch = make_batch(ds, 'source', 96, rng)
        batch.update(make_batch(ds, 'target', 96, rng))
        save(f'conet_{tag}', run_and_pack(m, batch, sizes_scalar(ds)))

    # ---------------- DTCDR (NeuMF base, dropout 0) ----------------
    torch.manual_seed(2022)
    m = DTCDR(base_config(embedding_size=D, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                          alpha=0.4), ds_b)
    rng = np.random.RandomState(23)
    batch = make_batch(ds_b, 'source', 96, rng)
    batch.update(make_batch(ds_b, 'target', 80, rng))
    extra = sizes_scalar(ds_b)
    extra['meta/alpha'] = 0.4
    save('dtcdr_neumf', run_and_pack(m, batch, extra))

    # ---------------- BiTGCF (drop_rate 0; concat and mean) ----------------
    # cross-check the COO builder that replaces the private-API build of bitgcf.py:92-116
    for dom in ('source', 'target'):
        im = ds_b.inter_matrix(form='coo', domain=dom).astype(np.float32)
        a = slow_norm_adj(im, ds_b.num_total_user, ds_b.num_total_item).to_dense()
        b = cdr_oracle.bitgcf_norm_adj(im.row, im.col, ds_b.num_total_user, ds_b.num_total_item).to_dense()
        assert torch.equal(a, b), 'norm-adj replacement differs from the literal bitgcf.py build'
    BiTGCF.get_norm_adj_mat = lambda self, inter, n_users=None, n_items=None: cdr_oracle.bitgcf_norm_adj(
        inter.row, inter.col, n_users, n_items)
    for way in ('concat', 'mean'):
        torch.manual_seed(2022)
        m = BiTGCF(base_config(embedding_size=32, n_layers=2, reg_weight=0.001, lambda_source=0.8, lambda_target=0.7,
                               drop_rate=0.0, connect_way=way), ds_b)
        rng = np.random.RandomState(29)
        batch = make_batch(ds_b, 'source', 96, rng)
        batch.update(make_batch(ds_b, 'target', 80, rng))
        extra = sizes_scalar(ds_b)
        for dom in ('source', 'target'):
            extra[f'edges/{dom}_row'], extra[f'edges/{dom}_col'] = ds_b.edges[dom]
        save(f'bitgcf_{way}', run_and_pack(m, batch, extra))


if __name__ == '__main__':
    main()
