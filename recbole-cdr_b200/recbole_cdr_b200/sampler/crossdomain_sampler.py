"""Negative samplers on the device -- mirror of reference sampler/crossdomain_sampler.py (uniform distribution).

``CrossDomainSourceSampler.sample_by_user_ids(user_ids, item_ids, num)`` keeps the reference's signature, output layout
(``num`` blocks of ``len(user_ids)``) and contract (uniform over the source domain's valid items, never an item the user
interacted with).  The draw and the rejection loop run in one CUDA kernel (``xdr_neg_sample_uniform``) over a CSR copy
of the used-item sets; ids can stay on the device for the following training step.  ``TargetDomainSampler`` is the
recbole ``Sampler`` the reference uses for the target domain (uniform over ``[1, item_num)``).

The popularity (alias-table) distribution of ``AbstractSampler._pop_sampling`` is not on the default path
(``neg_sampling: {uniform: 1}``, properties/overall.yaml:22-23) and raises NotImplementedError.
"""
import copy

import torch

from .._lib import call, cur_stream, ptr


def build_used_csr(user_ids, item_ids, n_users):
    """get_used_ids (crossdomain_sampler.py:229-250) as CSR: sorted, de-duplicated used items per user."""
    u = torch.as_tensor(user_ids, dtype=torch.int64).reshape(-1)
    i = torch.as_tensor(item_ids, dtype=torch.int64).reshape(-1)
    if u.numel():
        pairs = torch.unique(torch.stack([u, i], 1), dim=0)   # lexicographically sorted rows
    else:
        pairs = torch.zeros((0, 2), dtype=torch.int64)
    counts = torch.bincount(pairs[:, 0], minlength=n_users)
    rowptr = torch.zeros(n_users + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, pairs[:, 1].contiguous()


class _DeviceUniformSampler(object):
    def __init__(self, n_users, n_overlap, n_gap, n_valid, user_ids, item_ids, device, seed=2022, distribution='uniform'):
        self.set_distribution(distribution)
        self.device = torch.device(device)
        self.n_users, self.n_overlap, self.n_gap, self.n_valid = int(n_users), int(n_overlap), int(n_gap), int(n_valid)
        rowptr, col = build_used_csr(user_ids, item_ids, self.n_users)
        deg = rowptr[1:] - rowptr[:-1]
        if deg.numel() and int(deg.max()) >= self.n_valid:
            # same condition as crossdomain_sampler.py:243-249 (a user that interacted with every item)
            raise ValueError('Some users have interacted with all items, which we can not sample negative items for '
                             'them. Please set `user_inter_num_interval` to filter those users.')
        self.used_rowptr, self.used_col = rowptr.to(self.device), col.to(self.device)
        self.seed = int(seed)
        self._calls = 0
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.phase = None

    def set_distribution(self, distribution):
        if distribution == 'popularity':
            raise NotImplementedError('popularity sampling is outside the xdr hot-path scope (SURVEY.md section 2 row 9)')
        if distribution != 'uniform':
            raise NotImplementedError(f'The sampling distribution [{distribution}] is not implemented.')
        self.distribution = distribution

    def set_phase(self, phase):
        """Copy of this sampler bound to a phase (crossdomain_sampler.py:252-267); used ids are phase-independent here,
        as in the reference, whose get_used_ids reads the full source dataset for every phase."""
        new = copy.copy(self)
        new.phase = phase
        return new

    def sample_by_key_ids(self, key_ids, num, check=True):
        """[len(key_ids) * num] int64 on the device; block j holds the j-th draw of every key (np.tile layout)."""
        keys = torch.as_tensor(key_ids, dtype=torch.int64).to(self.device).reshape(-1).contiguous()
        out = torch.empty(keys.numel() * int(num), dtype=torch.int64, device=self.device)
        self._calls += 1
        call('xdr_neg_sample_uniform', ptr(keys), keys.numel(), int(num), ptr(self.used_rowptr), ptr(self.used_col),
             self.n_users, self.n_overlap, self.n_gap, self.n_valid, self.seed, self._calls & 0xFFFFFFFF, 1000, ptr(out),
             ptr(self._status), cur_stream())
        if check:
            self.check_status()
        return out

    def check_status(self):
        """Read (one device->host sync) and clear the status word the draws since the last check have OR-ed into: raises as
        the reference does for an unknown user id, and when a draw exhausted its attempts (its slot then holds an item the
        user HAS interacted with).  ``sample_by_key_ids(check=False)`` callers -- the device-resident epoch loops -- call it
        once per epoch."""
        st = int(self._status.item())
        if st:
            self._status.zero_()
            if st & 2:
                raise ValueError('user_id not exist.')
            raise ValueError('negative sampling exhausted its attempts for some user')


class CrossDomainSourceSampler(_DeviceUniformSampler):
    """Negative items for the SOURCE domain: candidates [1, n_ov_items) ++ [n_ov_items + n_tgt_only_items, n_total)
    (crossdomain_sampler.py:212-213)."""

    def __init__(self, phases, dataset, built_datasets=None, distribution='uniform', *, user_ids=None, item_ids=None,
                 device='cuda', seed=2022):
        self.phases = phases if isinstance(phases, list) else [phases]
        if user_ids is None:  # the reference's route: the source dataset's interaction columns
            src = dataset.source_domain_dataset
            user_ids, item_ids = src.inter_feat[src.uid_field], src.inter_feat[src.iid_field]
        self.overlapped_item_num = dataset.num_overlap_item
        self.target_only_item_num = dataset.num_target_only_item
        self.source_only_item_num = dataset.num_source_only_item
        self.total_item_num = dataset.num_total_item
        self.item_num = self.overlapped_item_num + self.source_only_item_num
        super().__init__(dataset.num_total_user, self.overlapped_item_num, self.target_only_item_num, self.item_num - 1,
                         user_ids, item_ids, device, seed, distribution)

    def sample_by_user_ids(self, user_ids, item_ids, num):
        return self.sample_by_key_ids(user_ids, num)


class TargetDomainSampler(_DeviceUniformSampler):
    """recbole.sampler.Sampler for the target domain [recbole-1.0.1]: uniform over [1, item_num)."""

    def __init__(self, n_users, item_num, user_ids, item_ids, device='cuda', seed=None, distribution='uniform'):
        # default key differs from the source sampler's: with one key and equal call counters the two domains of a BOTH-mode
        # step would draw from the same Philox stream position by position
        seed = (2022 ^ 0x5bd1e995) if seed is None else seed
        super().__init__(n_users, item_num, 0, item_num - 1, user_ids, item_ids, device, seed, distribution)

    def sample_by_user_ids(self, user_ids, item_ids, num):
        return self.sample_by_key_ids(user_ids, num)
