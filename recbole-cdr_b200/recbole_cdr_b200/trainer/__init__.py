from .trainer import CrossDomainTrainer, FusedStepRunner  # noqa: F401
from .graphed import GraphedTrainStep  # noqa: F401
