from .interaction import Interaction  # noqa: F401
from .idspace import IdSpace  # noqa: F401
