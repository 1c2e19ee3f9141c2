"""Import-only stand-in for dgl (see ../README.md)."""
from . import function, geometry, nn, utils  # noqa: F401


def _unavailable(*_a, **_k):
    raise NotImplementedError(
        "dgl is not installed; tpugan_b200 replaces gcn_lib.cubic_interpolation with "
        "tpugan_b200.interpolation.cubic_interpolation (see INTEGRATION.md)")


class DGLGraph:  # annotation target only
    pass


def graph(*a, **k):  # gcn_lib/graph_utils.py:61 evaluates `dgl.graph` in an annotation at import
    return _unavailable()
