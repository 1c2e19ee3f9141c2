"""Functional mini-emulation of the DGL calls made by gcn_lib/interpolation.py:16-123
(graph from (src, dst), add_edges, edges, local_scope, src/dst/e/n data, apply_edges with a
lambda, update_all with fn.src_mul_edge/fn.copy_e + fn.sum).  Edge order = insertion order;
messages are summed per destination in edge order.  Test infrastructure only."""
import contextlib
from types import SimpleNamespace

import torch

from . import function, geometry, nn, utils  # noqa: F401


class DGLGraph:
    def __init__(self, src, dst):
        self.src = src.clone().long()
        self.dst = dst.clone().long()
        self.srcdata, self.dstdata, self.edata, self.ndata = {}, {}, {}, {}

    def num_nodes(self):
        n = 0
        if self.src.numel():
            n = int(max(self.src.max(), self.dst.max())) + 1
        return n

    def add_edges(self, u, v):
        self.src = torch.cat([self.src, u.long()])
        self.dst = torch.cat([self.dst, v.long()])

    def edges(self):
        return self.src, self.dst

    @contextlib.contextmanager
    def local_scope(self):
        yield

    def apply_edges(self, func):
        out = func(SimpleNamespace(data=self.edata))
        self.edata.update(out)

    def update_all(self, message_func, reduce_func):
        kind, a, b, out_msg = message_func
        if kind == "src_mul_edge":
            m = self.srcdata[a][self.src] * self.edata[b]
        elif kind == "copy_e":
            m = self.edata[a]
        else:
            raise NotImplementedError(kind)
        rkind, msg_name, out_name = reduce_func
        assert rkind == "sum" and msg_name == out_msg
        n = self.srcdata["h"].shape[0] if "h" in self.srcdata else self.num_nodes()
        acc = torch.zeros((n,) + tuple(m.shape[1:]), dtype=m.dtype)
        acc.index_add_(0, self.dst, m)
        self.ndata[out_name] = acc
        self.dstdata[out_name] = acc


def graph(data, *a, **k):
    src, dst = data
    return DGLGraph(src, dst)
