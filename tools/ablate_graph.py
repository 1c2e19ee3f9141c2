#!/usr/bin/env python
"""Which op classes bound the DAG replay?  Captures the multi-stream CUDA graph of the step with one op class
stubbed out at a time (the stub returns the tensors a normal pass produced; no kernel runs) and reports the step
time.  GPU box only, tuning aid.   python tools/ablate_graph.py [lanes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
import hotpath_trace as ht  # noqa: E402
from tpugan_b200 import _lib  # noqa: E402

_lib.set_option("fps.sms_per_cloud", int(os.environ.get("FPS_SMS", "1")))
from tpugan_b200 import functional as _Fn  # noqa: E402
_Fn.csr_cache.prefetch_enabled = os.environ.get("CSR_PREFETCH", "1") == "1"
_Fn.knn_memo.enabled = os.environ.get("KNN_MEMO", "0") == "1"
_lib.set_option("fps.exclusive_sm", int(os.environ.get("FPS_EXCL", "1")))

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
only_full = len(sys.argv) > 2
doc = ht.load_schedule(os.path.join(ROOT, "tests", "golden", "fluid_step_schedule.json"), 8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


class Ops(ht.TorchCudaOps):
    stub = None
    record = None

    def _wrap(self, name, fn, *a):
        if self.record is not None:
            out = fn(*a)
            self.record.setdefault(name, []).append(out)
            return out
        if self.stub == name:
            self._pos[name] = self._pos.get(name, 0) + 1
            return self.cache[name][self._pos[name] - 1]
        return fn(*a)

    def new_step(self):
        super().new_step()
        self._pos = {}

    def knn(self, *a): return self._wrap("knn", super().knn, *a)
    def frnn(self, *a): return self._wrap("frnn", super().frnn, *a)
    def fps(self, *a): return self._wrap("fps", super().fps, *a)
    def gather(self, *a): return self._wrap("gather", super().gather, *a)
    def ball_query(self, *a): return self._wrap("ball_query", super().ball_query, *a)
    def group(self, *a): return self._wrap("group", super().group, *a)
    def group_bwd(self, *a): return self._wrap("group_bwd", super().group_bwd, *a)


ops = Ops("cuda")
rp = ht.TraceReplay(doc, ops, seed=1)
ops.record = {}
rp.run_step(lanes=lanes)  # same issue order as the measured runs (the cache is consumed by occurrence)
ops.cache, ops.record = ops.record, None
torch.cuda.synchronize()


def measure(stub):
    ops.stub = stub
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        rp.run_step(lanes=lanes)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        rp.run_step(lanes=lanes)
    ts = []
    for _ in range(8):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


base = measure(None)
print(f"lanes={lanes}  full step {base:.2f} ms")
for name in (() if only_full else ("fps", "knn", "group_bwd", "group", "ball_query", "gather", "frnn")):
    t = measure(name)
    print(f"  without {name:11s} {t:6.2f} ms   (-{base - t:.2f})")
