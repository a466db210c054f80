class TrainDataLoader(object):
    pass


class FullSortEvalDataLoader(object):
    pass
