from .interaction import Interaction  # noqa: F401
from .idspace import IdSpace  # noqa: F401
from .dataloader import CrossDomainDataloader, DomainTrainDataLoader, OverlapDataloader  # noqa: F401
from .device_pipeline import DeviceDomainData, DeviceDomainTrainDataLoader  # noqa: F401
