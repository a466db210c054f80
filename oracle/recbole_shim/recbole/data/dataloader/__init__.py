class NegSampleEvalDataLoader(object):
    pass
