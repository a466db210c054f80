class Dataset(object):
    pass
