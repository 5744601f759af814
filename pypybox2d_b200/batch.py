"""placeholder, replaced below"""
class BatchedWorld(object):
    pass
