"""Batch multiplexer of the hot path -- mirror of reference data/dataloader.py (`CrossDomainDataloader`, :55-180;
`OverlapDataloader`, :25-52) and of the recbole `TrainDataLoader` batch expansion it wraps [recbole-1.0.1].

Host-side control flow only: which fields a batch carries in each of the four states and how
the two domain loaders are stepped.  In-memory loaders over id arrays replace recbole's dataframe machinery; the
`Interaction` they yield has the reference's field names, dtypes (int64 ids, fp32 labels) and layouts:

* pointwise + ``neg_sampling: {uniform: 1}``: ``step = batch_size // 2`` positives per batch, batch = positives followed by
  the same users with sampled items, label column ``1..1, 0..0``;
* pairwise: ``step = batch_size`` positives and an extra ``NEG_PREFIX + iid`` column;
* OVERLAP: ``{'overlap': [b, 1]}`` drawn from a shuffled ``arange(num_overlap)`` INCLUDING the PAD id 0 (data/dataset.py:694-696).
"""
import torch

from ..utils.enum_type import CrossDomainDataLoaderState
from .interaction import Interaction


class DomainTrainDataLoader(object):
    """One domain's training loader: slices ``[pr, pr + step)`` of the (shuffled) interactions and expands negatives."""

    def __init__(self, uid_field, iid_field, users, items, batch_size, sampler, pairwise, label_field=None,
                 neg_prefix='neg_', shuffle=False, generator=None):
        self.uid_field, self.iid_field, self.label_field = uid_field, iid_field, label_field
        self.neg_iid_field = neg_prefix + iid_field
        self.users = torch.as_tensor(users, dtype=torch.int64)
        self.items = torch.as_tensor(items, dtype=torch.int64)
        self.batch_size, self.sampler, self.pairwise, self.shuffle = batch_size, sampler, pairwise, shuffle
        self.generator = generator
        self.step = batch_size if pairwise else max(batch_size // 2, 1)  # recbole: batch_num = batch_size // times
        self.pr = 0
        self.order = torch.arange(self.users.numel())

    @property
    def pr_end(self):
        return self.users.numel()

    def __len__(self):
        return -(-self.pr_end // self.step)

    def __iter__(self):
        if self.shuffle:
            self.order = torch.randperm(self.users.numel(), generator=self.generator)
        return self

    def __next__(self):
        if self.pr >= self.pr_end:
            self.pr = 0
            raise StopIteration()
        sel = self.order[self.pr:self.pr + self.step]
        self.pr += self.step
        u, i = self.users[sel], self.items[sel]
        neg = torch.as_tensor(self.sampler.sample_by_user_ids(u, i, 1)).to(torch.int64).cpu()
        if self.pairwise:
            return Interaction({self.uid_field: u, self.iid_field: i, self.neg_iid_field: neg})
        labels = torch.cat([torch.ones(u.numel()), torch.zeros(u.numel())])
        return Interaction({self.uid_field: torch.cat([u, u]), self.iid_field: torch.cat([i, neg]), self.label_field: labels})


class OverlapDataloader(object):
    """Batches of overlapped ids, shape [b, 1] (reference data/dataloader.py:25-52, data/dataset.py:686-696)."""

    def __init__(self, num_overlap, batch_size, shuffle=False, generator=None, field='overlap'):
        self.field, self.step, self.shuffle, self.generator = field, batch_size, shuffle, generator
        self.ids = torch.randperm(num_overlap, generator=generator).reshape(-1, 1)  # shuffled arange, PAD 0 included
        self.pr = 0

    @property
    def pr_end(self):
        return self.ids.shape[0]

    def __len__(self):
        return -(-self.pr_end // self.step)

    def __iter__(self):
        if self.shuffle:
            self.ids = self.ids[torch.randperm(self.ids.shape[0], generator=self.generator)]
        return self

    def __next__(self):
        if self.pr >= self.pr_end:
            self.pr = 0
            raise StopIteration()
        cur = self.ids[self.pr:self.pr + self.step]
        self.pr += self.step
        return Interaction({self.field: cur})


class CrossDomainDataloader(object):
    """4-state multiplexer over a source loader, a target loader and an overlap loader (data/dataloader.py:55-180).

    SOURCE / TARGET / OVERLAP delegate; BOTH yields ``target_batch.update(source_batch)``, the epoch ends with the
    TARGET loader and the source loader silently restarts when it runs out (:148-162)."""

    def __init__(self, source_dataloader, target_dataloader, overlap_dataloader):
        self.source_dataloader = source_dataloader
        self.target_dataloader = target_dataloader
        self.overlap_dataloader = overlap_dataloader
        self.state = CrossDomainDataLoaderState.BOTH

    def set_mode(self, state):
        if state not in set(CrossDomainDataLoaderState):
            raise NotImplementedError(f'Cross Domain data loader has no state named [{state}].')
        if self.source_dataloader.pr != 0 or self.target_dataloader.pr != 0:
            raise PermissionError('Cannot change dataloader\'s state within an epoch')
        self.state = state

    def __iter__(self):
        S = CrossDomainDataLoaderState
        if self.state == S.SOURCE:
            return self.source_dataloader.__iter__()
        if self.state == S.TARGET:
            return self.target_dataloader.__iter__()
        if self.state == S.OVERLAP:
            return self.overlap_dataloader.__iter__()
        self.source_dataloader.__iter__()
        self.target_dataloader.__iter__()
        return self

    def __next__(self):
        # only reached in the BOTH state (the other states iterate their own loader)
        if self.target_dataloader.pr >= self.target_dataloader.pr_end:
            self.target_dataloader.pr = 0
            self.source_dataloader.pr = 0
            raise StopIteration()
        try:
            source_data = self.source_dataloader.__next__()
        except StopIteration:
            source_data = self.source_dataloader.__next__()
        target_data = self.target_dataloader.__next__()
        target_data.update(source_data)
        return target_data

    def __len__(self):
        S = CrossDomainDataLoaderState
        if self.state == S.SOURCE:
            return len(self.source_dataloader)
        if self.state == S.OVERLAP:
            return len(self.overlap_dataloader)
        return len(self.target_dataloader)

    @property
    def pr_end(self):
        S = CrossDomainDataLoaderState
        if self.state == S.SOURCE:
            return self.source_dataloader.pr_end
        if self.state == S.OVERLAP:
            return self.overlap_dataloader.pr_end
        return self.target_dataloader.pr_end
