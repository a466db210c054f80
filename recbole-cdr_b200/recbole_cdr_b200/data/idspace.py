"""A0: the joint user/item id layout that every table on the hot path is indexed by.

Restates the layout produced by ``CrossDomainDataset.calculate_user_item_from_both_domain`` (reference
data/dataset.py:344-445) as plain integer arithmetic:

    0                               [PAD]                                   (dataset.py:391,435)
    [1, n_ov)                       ids present in both domains            (dataset.py:390)
    [n_ov, n_ov + n_tgt_only)       target-only ids                         (dataset.py:392-394)
    [n_ov + n_tgt_only, n_total)    source-only ids                         (dataset.py:395-396)

``n_ov`` counts the PAD row (``num_overlap_* = len(overlap) + 1``, dataset.py:384,428), so ``n_ov == 1`` means "no
overlap".  Every table is allocated with ``n_total`` rows (emcdr.py:67-71): source tables have dead rows
``[n_ov, target_num)`` and target tables dead rows ``[target_num, n_total)``.
"""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class IdSpace:
    n_overlap: int      # incl. PAD
    n_target_only: int
    n_source_only: int

    @property
    def n_total(self) -> int:
        return self.n_overlap + self.n_target_only + self.n_source_only

    @property
    def target_num(self) -> int:  # crossdomain_recommender.py:35-36
        return self.n_overlap + self.n_target_only

    @property
    def source_num(self) -> int:  # crossdomain_recommender.py:28-29
        return self.n_overlap + self.n_source_only

    def n_valid(self, domain: str) -> int:
        """Number of real (non-PAD) ids a domain can emit."""
        return (self.n_overlap - 1) + (self.n_source_only if domain == 'source' else self.n_target_only)

    def compact_to_joint(self, compact: torch.Tensor, domain: str) -> torch.Tensor:
        """Map k in [0, n_valid(domain)) to the k-th valid joint id of the domain (the candidate list of the source
        sampler, sampler/crossdomain_sampler.py:212-213: [1, n_ov) ++ [target_num, n_total))."""
        k = compact + 1
        if domain == 'source':
            return torch.where(k < self.n_overlap, k, k + self.n_target_only)
        return k

    def joint_to_compact(self, joint: torch.Tensor, domain: str) -> torch.Tensor:
        if domain == 'source':
            return torch.where(joint < self.n_overlap, joint, joint - self.n_target_only) - 1
        return joint - 1

    def is_valid(self, joint: torch.Tensor, domain: str) -> torch.Tensor:
        if domain == 'source':
            return ((joint >= 1) & (joint < self.n_overlap)) | ((joint >= self.target_num) & (joint < self.n_total))
        return (joint >= 1) & (joint < self.target_num)

    def valid_ids_numpy(self, domain: str) -> np.ndarray:
        if domain == 'source':
            return np.concatenate([np.arange(1, self.n_overlap), np.arange(self.target_num, self.n_total)])
        return np.arange(1, self.target_num)
