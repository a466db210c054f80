"""Synthetic workloads of the shapes BASELINE.json names: id layout + seeded batches.

There is no network for datasets, so benchmarks and large-size tests draw ``(user, item, neg)`` triples and
pointwise ``(user, item, label)`` rows uniformly (or Zipf) over each domain's *valid* id ranges of the joint layout.
"""
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from .idspace import IdSpace


class _Domain:
    def __init__(self, prefix, n_users, n_items):
        self.uid_field = f'{prefix}_user_id'
        self.iid_field = f'{prefix}_item_id'
        self.label_field = f'{prefix}_label'
        self._num = {self.uid_field: n_users, self.iid_field: n_items}

    def num(self, field):
        return self._num[field]


@dataclass
class SyntheticCrossDomainDataset:
    """The dataset attributes the model base class reads (reference model/crossdomain_recommender.py:24-45)."""
    users: IdSpace
    items: IdSpace

    def __post_init__(self):
        self.num_overlap_user, self.num_overlap_item = self.users.n_overlap, self.items.n_overlap
        self.num_target_only_user, self.num_source_only_user = self.users.n_target_only, self.users.n_source_only
        self.num_target_only_item, self.num_source_only_item = self.items.n_target_only, self.items.n_source_only
        self.num_total_user, self.num_total_item = self.users.n_total, self.items.n_total
        self.source_domain_dataset = _Domain('source', self.users.source_num, self.items.source_num)
        self.target_domain_dataset = _Domain('target', self.users.target_num, self.items.target_num)
        self.overlap_id_field = 'overlap'


def emcdr_scale(scale: int) -> SyntheticCrossDomainDataset:
    """BASELINE configs #2 (scale=1_000_000) and #5 (10_000_000): user-overlap scenario, half of each domain's users
    overlapped, items disjoint."""
    half = scale // 2
    return SyntheticCrossDomainDataset(IdSpace(half + 1, half, half), IdSpace(1, scale, scale))


def draw_ids(space: IdSpace, domain: str, n: int, gen: torch.Generator, device, zipf: Optional[float] = None):
    nv = space.n_valid(domain)
    if zipf is None:
        k = torch.randint(0, nv, (n,), generator=gen, device=device, dtype=torch.int64)
    else:  # inverse-CDF approximation of a bounded Zipf(s): rank ~ u^(-1/(s-1)), popular ranks first
        u = torch.rand(n, generator=gen, device=device, dtype=torch.float64).clamp_min(1e-12)
        k = (u.pow(-1.0 / max(zipf - 1.0, 1e-3)) - 1.0).clamp(0, nv - 1).to(torch.int64)
    return space.compact_to_joint(k, domain)


def make_batch(ds: SyntheticCrossDomainDataset, domain: str, batch: int, seed: int, device='cpu', pairwise=True,
               zipf_items: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """One per-domain batch with the reference's field names; seed convention: Generator().manual_seed(1 + step)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    out = {f'{domain}_user_id': draw_ids(ds.users, domain, batch, gen, device),
           f'{domain}_item_id': draw_ids(ds.items, domain, batch, gen, device, zipf_items)}
    if pairwise:
        out[f'neg_{domain}_item_id'] = draw_ids(ds.items, domain, batch, gen, device, zipf_items)
    else:
        half = batch // 2  # recbole pointwise batches: positives then sampled negatives, labels 1..1,0..0
        out[f'{domain}_label'] = torch.cat([torch.ones(half), torch.zeros(batch - half)]).to(device)
    return out
