"""A stand-in for CrossDomainDataset exposing exactly the attributes the model base class reads
(reference model/crossdomain_recommender.py:24-45) on the joint id layout of data/dataset.py:344-445."""
import numpy as np
import scipy.sparse as sp
import torch


class _Domain:
    def __init__(self, prefix, n_users, n_items):
        self.uid_field = f'{prefix}_user_id'
        self.iid_field = f'{prefix}_item_id'
        self.label_field = f'{prefix}_label'
        self._num = {self.uid_field: n_users, self.iid_field: n_items}

    def num(self, field):
        return self._num[field]


class FakeDataset:
    def __init__(self, n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, edges=None):
        self.num_overlap_user, self.num_overlap_item = n_ov_u, n_ov_i
        self.num_target_only_user, self.num_source_only_user = n_tgt_u, n_src_u
        self.num_target_only_item, self.num_source_only_item = n_tgt_i, n_src_i
        self.num_total_user = n_ov_u + n_tgt_u + n_src_u
        self.num_total_item = n_ov_i + n_tgt_i + n_src_i
        self.source_domain_dataset = _Domain('source', n_ov_u + n_src_u, n_ov_i + n_src_i)
        self.target_domain_dataset = _Domain('target', n_ov_u + n_tgt_u, n_ov_i + n_tgt_i)
        self.overlap_id_field = 'overlap'
        self.edges = edges or {}

    @classmethod
    def from_golden(cls, g, edges=None):
        return cls(g.meta('n_ov_u'), g.meta('n_tgt_u'), g.meta('n_src_u'), g.meta('n_ov_i'), g.meta('n_tgt_i'),
                   g.meta('n_src_i'), edges)

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


def base_config(device='cuda', **kw):
    cfg = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': device}
    cfg.update(kw)
    return cfg


def make_batch(ds, domain, B, rng, pairwise=False, zipf=None):
    users, items = ds.valid_ids(domain)

    def draw(pool):
        if zipf is None:
            return rng.choice(pool, B)
        r = np.minimum(rng.zipf(zipf, B) - 1, len(pool) - 1)
        return pool[r]

    b = {f'{domain}_user_id': torch.from_numpy(draw(users)).long(),
         f'{domain}_item_id': torch.from_numpy(draw(items)).long()}
    if pairwise:
        b[f'neg_{domain}_item_id'] = torch.from_numpy(draw(items)).long()
    else:
        b[f'{domain}_label'] = torch.from_numpy((rng.rand(B) < 0.5).astype(np.float32))
    return b
