"""The batch dict of the hot path: a mirror of ``recbole.data.interaction.Interaction`` [recbole-1.0.1] restricted to
what RecBole-CDR's dataloaders and models use (reference data/dataloader.py:18,155-162,229)."""
import torch


class Interaction(object):
    """dict of equal-length tensors; ``[str] -> Tensor``, ``.to(device)``, ``.update(other)``, ``len()``."""

    def __init__(self, interaction):
        self.interaction = dict()
        for k, v in interaction.items():
            if not isinstance(v, torch.Tensor):
                v = torch.as_tensor(v)
            self.interaction[k] = v
        self.length = -1
        for v in self.interaction.values():
            self.length = max(self.length, v.shape[0])

    def __iter__(self):
        return iter(self.interaction)

    def __contains__(self, item):
        return item in self.interaction

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.interaction[index]
        return Interaction({k: v[index] for k, v in self.interaction.items()})

    def __setitem__(self, key, value):
        self.interaction[key] = value
        self.length = max(self.length, value.shape[0])

    def __len__(self):
        return self.length

    @property
    def columns(self):
        return list(self.interaction.keys())

    def to(self, device, selected_field=None, non_blocking=False):
        out = {}
        for k, v in self.interaction.items():
            if selected_field is None or k in selected_field:
                out[k] = v.to(device, non_blocking=non_blocking)
            else:
                out[k] = v
        return Interaction(out)

    def cpu(self):
        return self.to('cpu')

    def pin_memory(self):
        return Interaction({k: v.pin_memory() for k, v in self.interaction.items()})

    def update(self, new_inter):
        """Merge the fields of another Interaction in place (BOTH mode: target batch .update(source batch),
        reference data/dataloader.py:155-162; the two halves may differ in length on the last batch)."""
        for k in new_inter.interaction:
            self.interaction[k] = new_inter.interaction[k]
            self.length = max(self.length, new_inter.interaction[k].shape[0])

    def add_prefix(self, prefix):
        self.interaction = {prefix + k: v for k, v in self.interaction.items()}

    def repeat(self, sizes):
        return Interaction({k: v.repeat([sizes] + [1] * (v.dim() - 1)) for k, v in self.interaction.items()})

    def __str__(self):
        info = [f'The batch_size of interaction: {self.length}']
        for k, v in self.interaction.items():
            info.append(f'    {k}, {tuple(v.shape)}, {v.device.type}, {v.dtype}')
        return '\n'.join(info) + '\n'

    __repr__ = __str__
