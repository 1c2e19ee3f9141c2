def src_mul_edge(src, edge, out):
    return ("src_mul_edge", src, edge, out)


u_mul_e = src_mul_edge


def copy_e(edge, out):
    return ("copy_e", edge, None, out)


def sum(msg, out):  # noqa: A001
    return ("sum", msg, out)
