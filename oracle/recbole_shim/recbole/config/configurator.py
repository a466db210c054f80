class Config(object):
    """placeholder: only subclassed by recbole_cdr.config.CDRConfig, never instantiated by the oracle"""
    pass
