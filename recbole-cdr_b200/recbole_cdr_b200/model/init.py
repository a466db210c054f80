"""``recbole.model.init.xavier_normal_initialization`` [recbole-1.0.1]: Embedding/Linear weight <- xavier_normal_,
Linear bias <- 0 (call sites emcdr.py:84, conet.py:89, bitgcf.py:89, dtcdr.py:104, cmf.py:51)."""
import torch.nn as nn
from torch.nn.init import constant_, xavier_normal_


def xavier_normal_initialization(module):
    if isinstance(module, nn.Embedding):
        xavier_normal_(module.weight.data)
    elif isinstance(module, nn.Linear):
        xavier_normal_(module.weight.data)
        if module.bias is not None:
            constant_(module.bias.data, 0)
