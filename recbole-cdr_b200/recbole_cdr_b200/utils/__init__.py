from .enum_type import ModelType, InputType, CrossDomainDataLoaderState, train_mode2state  # noqa: F401
from .utils import get_model, get_trainer  # noqa: F401
