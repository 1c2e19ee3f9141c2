def farthest_point_sampler(*a, **k):
    raise NotImplementedError
