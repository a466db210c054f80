from .trainer import CrossDomainTrainer, DCDCSRTrainer, FusedStepRunner  # noqa: F401
from .graphed import GraphedTrainStep  # noqa: F401
from .row_optim import RowSparseOptimizer  # noqa: F401
