"""Device-resident batch pipeline: positives are permuted and negatives are drawn ON THE
GPU, producing the `[K, rows, B]` id blocks that `xdr_train_steps` consumes -- no host loop, no PCIe on the step path.

Replaces, for one domain and one epoch, the host side of recbole `TrainDataLoader.__next__` + `_neg_sampling`
[recbole-1.0.1] as wrapped by reference data/dataloader.py:72-76 and fed by sampler/crossdomain_sampler.py:269-290:
shuffle the interactions, slice `step` positives per batch, draw one negative per positive that the user has not
interacted with.  Field semantics are the reference's: pairwise -> (user, item, neg_item); pointwise -> positives followed
by (user, sampled item) with labels 1..1, 0..0.
"""
import torch


class DeviceDomainData(object):
    """One domain's training interactions resident in HBM + its device sampler."""

    def __init__(self, users, items, sampler, device='cuda'):
        self.device = torch.device(device)
        self.users = torch.as_tensor(users, dtype=torch.int64).to(self.device)
        self.items = torch.as_tensor(items, dtype=torch.int64).to(self.device)
        self.sampler = sampler

    def __len__(self):
        return self.users.numel()

    def epoch_blocks(self, batch_size, steps_per_block, pairwise=True, shuffle=True, generator=None, drop_last=True):
        """Yield `(ids [K, rows, B], label [K, B] or None)` blocks covering one epoch (K <= steps_per_block).

        pairwise: rows = (user, item, neg_item), B = batch_size positives per step.
        pointwise: rows = (user, item), B = batch_size with batch_size // 2 positives followed by their negatives.
        The ragged tail that does not fill a whole step is dropped when ``drop_last`` (the persistent kernel wants equal
        steps); otherwise it is returned as a final block of one shorter step, rounded down to a multiple of 4 interactions
        (the kernel's id tiles are 16-byte granular: at most 3 interactions of an epoch are left out)."""
        n = self.users.numel()
        perm = torch.randperm(n, device=self.device, generator=generator) if shuffle else torch.arange(n, device=self.device)
        pos = batch_size if pairwise else max(batch_size // 2, 1)
        n_steps = n // pos
        for s0 in range(0, n_steps, steps_per_block):
            k = min(steps_per_block, n_steps - s0)
            sel = perm[s0 * pos:(s0 + k) * pos]
            yield self._block(sel, k, pos, pairwise)
        tail = (n - n_steps * pos) // 4 * 4
        if not drop_last and tail > 0:
            sel = perm[n_steps * pos:n_steps * pos + tail]
            yield self._block(sel, 1, tail, pairwise)

    def _block(self, sel, k, pos, pairwise):
        u = self.users[sel]
        i = self.items[sel]
        neg = self.sampler.sample_by_key_ids(u, 1, check=False)
        u, i, neg = u.view(k, pos), i.view(k, pos), neg.view(k, pos)
        if pairwise:
            return torch.stack([u, i, neg], dim=1).contiguous(), None
        ids = torch.stack([torch.cat([u, u], dim=1), torch.cat([i, neg], dim=1)], dim=1).contiguous()
        # the persistent kernel walks ids and labels with ONE step stride: the label rows live in a [K, 2, B] block too
        label = torch.zeros((k, 2, 2 * pos), dtype=torch.float32, device=ids.device)[:, 0]
        label[:, :pos] = 1.0
        return ids, label


class DeviceDomainTrainDataLoader(object):
    """The device-resident twin of ``dataloader.DomainTrainDataLoader``: same fields, slicing, pointwise / pairwise expansion and
    ragged last batch, but the interactions live in HBM, the per-epoch shuffle is a device ``randperm`` and the negatives
    come from the device sampler -- every ``Interaction`` it yields is already on the device and no host array is touched.
    It plugs into ``CrossDomainDataloader`` (all four states) exactly like the host loader, so ANY model trains from it
    through the unchanged trainer loop ("Interaction-compatible device tensors")."""

    def __init__(self, uid_field, iid_field, users, items, batch_size, sampler, pairwise, label_field=None, neg_prefix='neg_',
                 shuffle=False, generator=None, device='cuda'):
        from .interaction import Interaction
        self._Interaction = Interaction
        self.device = torch.device(device)
        self.uid_field, self.iid_field, self.label_field = uid_field, iid_field, label_field
        self.neg_iid_field = neg_prefix + iid_field
        self.users = torch.as_tensor(users, dtype=torch.int64).to(self.device)
        self.items = torch.as_tensor(items, dtype=torch.int64).to(self.device)
        self.batch_size, self.sampler, self.pairwise, self.shuffle = batch_size, sampler, pairwise, shuffle
        self.generator = generator      # a generator on `device` (or None)
        self.step = batch_size if pairwise else max(batch_size // 2, 1)
        self.pr = 0
        self.order = torch.arange(self.users.numel(), device=self.device)

    @property
    def pr_end(self):
        return self.users.numel()

    def __len__(self):
        return -(-self.pr_end // self.step)

    def __iter__(self):
        if self.shuffle:
            self.order = torch.randperm(self.users.numel(), device=self.device, generator=self.generator)
        return self

    def __next__(self):
        if self.pr >= self.pr_end:
            self.pr = 0
            raise StopIteration()
        sel = self.order[self.pr:self.pr + self.step]
        self.pr += self.step
        u, i = self.users[sel], self.items[sel]
        neg = self.sampler.sample_by_key_ids(u, 1, check=False)
        if self.pairwise:
            return self._Interaction({self.uid_field: u, self.iid_field: i, self.neg_iid_field: neg})
        labels = torch.cat([torch.ones(u.numel(), device=self.device), torch.zeros(u.numel(), device=self.device)])
        return self._Interaction({self.uid_field: torch.cat([u, u]), self.iid_field: torch.cat([i, neg]),
                                  self.label_field: labels})
