#!/usr/bin/env python
"""Timeline of the DAG replay inside its CUDA graph: a 1-thread kernel stamps %globaltimer on the call's stream
before and after every call.  Prints per-op busy spans, concurrency over time and the largest gaps.
GPU box only, tuning aid.   python tools/timeline.py [lanes]"""
import collections
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
import hotpath_trace as ht  # noqa: E402
from tpugan_b200 import _lib  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = _lib.load()
_lib.set_option("fps.sms_per_cloud", 1)
from tpugan_b200 import functional as _Fn  # noqa: E402
_Fn.csr_cache.prefetch_enabled = os.environ.get("CSR_PREFETCH", "1") == "1"
_Fn.knn_memo.enabled = os.environ.get("KNN_MEMO", "0") == "1"
_lib.set_option("fps.exclusive_sm", 1)
ts_fn = lib.tpg_debug_timestamp
ts_fn.restype = ctypes.c_int
ts_fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
doc = ht.load_schedule(os.path.join(ROOT, "tests", "golden", "fluid_step_schedule.json"), 8)
ncalls = len(doc["calls"])
stamps = torch.zeros(2 * ncalls + 2, dtype=torch.int64, device="cuda")


class Ops(ht.TorchCudaOps):
    n = 0

    def new_step(self):
        super().new_step()
        self.n = 0

    def tick(self):
        ts_fn(stamps.data_ptr() + 8 * self.n, torch.cuda.current_stream().cuda_stream)
        self.n += 1
        return self.n


ops = Ops("cuda")
rp = ht.TraceReplay(doc, ops, seed=1)
rp.timers = collections.defaultdict(list)
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    rp.run_step(lanes=lanes)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    rp.run_step(lanes=lanes)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.fill_(1)
    g.replay()
torch.cuda.synchronize()
t = stamps.cpu().numpy()[: 2 * ncalls].reshape(ncalls, 2).astype(np.float64)
order = rp._lane_plan[2]
if order is not None:  # stamps are taken in issue order
    tt = np.empty_like(t)
    tt[np.asarray(order)] = t
    t = tt
assert (t > 0).all(), 'missing stamps'
t0 = t[:, 0].min()
t = (t - t0) / 1e3  # us
end = t[:, 1].max()
print(f"lanes={lanes} step span {end / 1e3:.2f} ms (with {2 * ncalls} stamp kernels)")
plan = rp._lane_plan[1]
ops_ = [c["op"] for c in doc["calls"]]
# concurrency profile
edges = sorted([(a, 1) for a in t[:, 0]] + [(b, -1) for b in t[:, 1]])
cur, last, hist = 0, 0.0, collections.Counter()
for x, d in edges:
    hist[cur] += x - last
    last = x
    cur += d
print("time (us) with k calls in flight:", {k: round(v) for k, v in sorted(hist.items())})
by = collections.defaultdict(float)
for n in range(ncalls):
    by[ops_[n]] += t[n, 1] - t[n, 0]
print("summed in-graph call spans (us):", {k: round(v) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
# coarse Gantt: 100-us buckets, which op classes are active
B = 250.0
nb = int(end / B) + 1
print(f"buckets of {B:.0f} us: ops active (count)")
for bi in range(nb):
    lo, hi = bi * B, (bi + 1) * B
    act = collections.Counter()
    for n in range(ncalls):
        if t[n, 0] < hi and t[n, 1] > lo:
            act[ops_[n]] += 1
    print(f"  {lo / 1e3:5.2f} ms  " + " ".join(f"{k}:{v}" for k, v in sorted(act.items())))
if len(sys.argv) > 2:
    for n in range(ncalls):
        i = doc["calls"][n]["in"]
        shp = [v.get("shape") for k, v in i.items() if isinstance(v, dict) and "shape" in v][:1]
        print(f"{n:4d} lane {plan[n]:2d} {ops_[n]:11s} {t[n, 0]:8.1f} -> {t[n, 1]:8.1f}  ({t[n, 1] - t[n, 0]:7.1f})  {shp} deps {sorted(rp._deps(n))[-4:]}")
