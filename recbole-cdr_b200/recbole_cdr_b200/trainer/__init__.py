from .trainer import CrossDomainTrainer, FusedStepRunner  # noqa: F401
