"""Reflection helpers with the reference's naming rule (reference utils/utils.py:16-59)."""
import importlib
import importlib.util

from .enum_type import ModelType


def get_model(model_name):
    """``recbole_cdr_b200.model.cross_domain_recommender.<name.lower()>.<name>``; ValueError for unknown names
    (same rule and error as reference utils/utils.py:16-40)."""
    module_path = '.'.join(['recbole_cdr_b200.model.cross_domain_recommender', model_name.lower()])
    if importlib.util.find_spec(module_path) is None:
        raise ValueError('`model_name` [{}] is not the name of an existing model.'.format(model_name))
    return getattr(importlib.import_module(module_path), model_name)


def get_trainer(model_type, model_name):
    """``<name>Trainer`` if it exists, else ``CrossDomainTrainer`` (reference utils/utils.py:43-59)."""
    mod = importlib.import_module('recbole_cdr_b200.trainer')
    try:
        return getattr(mod, model_name + 'Trainer')
    except AttributeError:
        if model_type == ModelType.CROSSDOMAIN:
            return getattr(mod, 'CrossDomainTrainer')
        raise
