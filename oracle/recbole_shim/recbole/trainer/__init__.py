class Trainer(object):
    """placeholder: parent of recbole_cdr.trainer.CrossDomainTrainer; not executed by the oracle"""
    def __init__(self, config, model):
        self.config = config
        self.model = model


class HyperTuning(object):
    pass
