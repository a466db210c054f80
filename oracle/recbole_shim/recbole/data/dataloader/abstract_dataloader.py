class AbstractDataLoader(object):
    pass
