"""[recbole-1.0.1] xavier_normal_initialization: Embedding/Linear weight <- xavier_normal_, Linear bias <- 0."""
import torch.nn as nn
from torch.nn.init import xavier_normal_, constant_


def xavier_normal_initialization(module):
    if isinstance(module, nn.Embedding):
        xavier_normal_(module.weight.data)
    elif isinstance(module, nn.Linear):
        xavier_normal_(module.weight.data)
        if module.bias is not None:
            constant_(module.bias.data, 0)
