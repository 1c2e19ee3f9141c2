def farthest_point_sampler(*_a, **_k):
    raise NotImplementedError("dgl.geometry is an import-only stand-in; use tpugan_b200.sampling")
