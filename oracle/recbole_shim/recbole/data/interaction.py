class Interaction(dict):
    """placeholder: the oracle passes plain dicts of tensors, which is all the models index"""
    pass
