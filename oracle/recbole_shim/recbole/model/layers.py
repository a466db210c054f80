"""[recbole-1.0.1] MLPLayers: per layer Dropout(p) -> Linear -> [BatchNorm1d] -> activation (after EVERY layer)."""
import torch.nn as nn
from torch.nn.init import normal_


def activation_layer(name, emb_dim=None):
    if name is None:
        return None
    name = name.lower()
    if name == 'sigmoid':
        return nn.Sigmoid()
    if name == 'tanh':
        return nn.Tanh()
    if name == 'relu':
        return nn.ReLU()
    if name == 'leakyrelu':
        return nn.LeakyReLU()
    if name == 'none':
        return None
    raise NotImplementedError(name)


class MLPLayers(nn.Module):
    def __init__(self, layers, dropout=0., activation='relu', bn=False, init_method=None):
        super().__init__()
        self.layers = layers
        self.dropout = dropout
        self.activation = activation
        self.use_bn = bn
        self.init_method = init_method
        mods = []
        for d_in, d_out in zip(self.layers[:-1], self.layers[1:]):
            mods.append(nn.Dropout(p=self.dropout))
            mods.append(nn.Linear(d_in, d_out))
            if self.use_bn:
                mods.append(nn.BatchNorm1d(num_features=d_out))
            act = activation_layer(self.activation, d_out)
            if act is not None:
                mods.append(act)
        self.mlp_layers = nn.Sequential(*mods)
        if self.init_method is not None:
            self.apply(self.init_weights)

    def init_weights(self, module):
        if isinstance(module, nn.Linear):
            if self.init_method == 'norm':
                normal_(module.weight.data, 0, 0.01)
            if module.bias is not None:
                module.bias.data.fill_(0.0)

    def forward(self, input_feature):
        return self.mlp_layers(input_feature)
