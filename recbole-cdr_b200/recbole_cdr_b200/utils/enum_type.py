"""Enums of the RecBole-CDR plugin API (reference utils/enum_type.py:18-45 and recbole.utils.InputType)."""
from enum import Enum


class ModelType(Enum):
    CROSSDOMAIN = 1


class InputType(Enum):
    POINTWISE = 1
    PAIRWISE = 2
    LISTWISE = 3


class CrossDomainDataLoaderState(Enum):
    BOTH = 1
    SOURCE = 2
    TARGET = 3
    OVERLAP = 4


train_mode2state = {name: getattr(CrossDomainDataLoaderState, name) for name in ('BOTH', 'SOURCE', 'TARGET', 'OVERLAP')}
