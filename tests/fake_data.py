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


class _DomainF4(_Domain):
    """Domain part that also carries the interaction list (``inter_feat``) some models read (sscdr.py:41,72-86)."""

    def __init__(self, prefix, n_users, n_items, rows, cols):
        super().__init__(prefix, n_users, n_items)
        self.inter_feat = {self.uid_field: torch.from_numpy(np.asarray(rows)).long(),
                           self.iid_field: torch.from_numpy(np.asarray(cols)).long()}

    def history(self, row_num, row):
        """Padded history matrix, zero values, lengths -- the algorithm of reference data/dataset.py:188-249
        (rows filled in interaction order, 0 = padding)."""
        u, i = self.inter_feat[self.uid_field].numpy(), self.inter_feat[self.iid_field].numpy()
        row_ids, col_ids = (u, i) if row == 'user' else (i, u)
        lens = np.bincount(row_ids, minlength=row_num).astype(np.int64)
        mat = np.zeros((row_num, max(int(lens.max()), 1) if len(row_ids) else 1), dtype=np.int64)
        fill = np.zeros(row_num, dtype=np.int64)
        for r, c in zip(row_ids, col_ids):
            mat[r, fill[r]] = c
            fill[r] += 1
        return torch.LongTensor(mat), torch.zeros(mat.shape), torch.LongTensor(lens)


class FakeDatasetF4(FakeDataset):
    """FakeDataset + what SSCDR / NATR / DCDCSR read: per-domain interaction lists and history matrices."""

    def __init__(self, n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, edges):
        super().__init__(n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, edges)
        self.source_domain_dataset = _DomainF4('source', n_ov_u + n_src_u, n_ov_i + n_src_i, *edges['source'])
        self.target_domain_dataset = _DomainF4('target', n_ov_u + n_tgt_u, n_ov_i + n_tgt_i, *edges['target'])

    @classmethod
    def random(cls, n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, seed=0, per_user=4):
        tmp = FakeDataset(n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i)
        rng = np.random.RandomState(seed)
        edges = {}
        for dom in ('source', 'target'):
            users, items = tmp.valid_ids(dom)
            r = np.repeat(users, per_user)
            c = rng.choice(items, size=r.shape[0])
            e = np.unique(np.stack([r, c], 1), axis=0)
            e = e[rng.permutation(len(e))]          # interaction order is not sorted in a real dataset
            edges[dom] = (e[:, 0].copy(), e[:, 1].copy())
        return cls(n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, edges)

    @classmethod
    def from_golden(cls, g, edges=None):
        edges = {dom: (g.z[f'edges/{dom}_row'], g.z[f'edges/{dom}_col']) for dom in ('source', 'target')}
        return cls(g.meta('n_ov_u'), g.meta('n_tgt_u'), g.meta('n_src_u'), g.meta('n_ov_i'), g.meta('n_tgt_i'),
                   g.meta('n_src_i'), edges)

    def history_item_matrix(self, value_field=None, domain='source'):
        part = self.source_domain_dataset if domain == 'source' else self.target_domain_dataset
        return part.history(self.num_total_user, 'user')

    def history_user_matrix(self, value_field=None, domain='source'):
        part = self.source_domain_dataset if domain == 'source' else self.target_domain_dataset
        return part.history(self.num_total_item, 'item')
