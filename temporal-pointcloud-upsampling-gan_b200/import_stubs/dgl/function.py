def _unavailable(*_a, **_k):
    raise NotImplementedError("dgl.function is an import-only stand-in (tpugan_b200/import_stubs)")


src_mul_edge = copy_e = copy_u = u_mul_e = sum = max = mean = _unavailable  # noqa: A001
