"""``recbole.model.layers.MLPLayers`` [recbole-1.0.1] as a parameter container with the reference's ``state_dict``
layout (``mlp_layers.{1,4,...}.weight``): per layer Dropout(p) -> Linear -> activation after EVERY layer.  ``forward``
runs each Linear+activation as one xdr dense kernel (dtcdr.py:61-67,121-124)."""
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, ops


class MLPLayers(nn.Module):
    def __init__(self, layers, dropout=0., activation='relu', bn=False, init_method=None):
        super().__init__()
        if bn:
            raise NotImplementedError('MLPLayers(bn=True) is not on the RecBole-CDR hot path')
        self.layers = layers
        self.dropout = dropout
        self.activation = activation
        self.use_bn = bn
        self.init_method = init_method
        self._act = _lib.ACT_BY_NAME[activation.lower() if isinstance(activation, str) else activation]
        mods = []
        for d_in, d_out in zip(self.layers[:-1], self.layers[1:]):
            mods.append(nn.Dropout(p=self.dropout))
            mods.append(nn.Linear(d_in, d_out))
            if self._act == _lib.ACT_RELU:
                mods.append(nn.ReLU())
            elif self._act == _lib.ACT_TANH:
                mods.append(nn.Tanh())
            elif self._act == _lib.ACT_SIGMOID:
                mods.append(nn.Sigmoid())
        self.mlp_layers = nn.Sequential(*mods)

    def forward(self, x):
        for m in self.mlp_layers:
            if isinstance(m, nn.Dropout):
                if self.training and m.p > 0:
                    x = F.dropout(x, m.p, True)
            elif isinstance(m, nn.Linear):
                x = ops.dense(x, m.weight, m.bias, self._act)
        return x
