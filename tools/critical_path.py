#!/usr/bin/env python
"""Critical path of the recorded step DAG with measured per-call device times (GPU box).
Per-call time = CUDA events around the call with a device synchronize before it (no launch-gap inflation).
    python tools/critical_path.py [fluid|action] [batch]"""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
import hotpath_trace as ht  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fluid"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
doc = ht.load_schedule(os.path.join(ROOT, "tests", "golden", f"{name}_step_schedule.json"), batch)


class SyncOps(ht.TorchCudaOps):
    def tick(self):
        torch.cuda.synchronize()
        return super().tick()


ops = SyncOps("cuda")
rp = ht.TraceReplay(doc, ops, seed=1)
for _ in range(2):
    rp.run_step()
reps = 5
acc = np.zeros(len(doc["calls"]))
for _ in range(reps):
    rp.timers = {}
    rp.run_step()
    torch.cuda.synchronize()
    cur = collections.Counter()
    for n, c in enumerate(doc["calls"]):
        a, b = rp.timers[c["op"]][cur[c["op"]]]
        cur[c["op"]] += 1
        acc[n] += a.elapsed_time(b) * 1e3
rp.timers = None
dur = acc / reps
calls = doc["calls"]
finish = np.zeros(len(calls))
pred = [-1] * len(calls)
for n in range(len(calls)):
    best, arg = 0.0, -1
    for d in rp._deps(n):
        if finish[d] > best:
            best, arg = finish[d], d
    finish[n] = best + dur[n]
    pred[n] = arg
end = int(np.argmax(finish))
path = []
while end >= 0:
    path.append(end)
    end = pred[end]
path.reverse()
print(f"sum of call times {dur.sum() / 1e3:.2f} ms; critical path {finish.max() / 1e3:.2f} ms over {len(path)} calls")
by = collections.defaultdict(float)
for n in path:
    by[calls[n]["op"]] += dur[n]
print("critical path by op:", {k: round(v / 1e3, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
tot = collections.defaultdict(float)
for n, c in enumerate(calls):
    tot[c["op"]] += dur[n]
print("all calls by op:   ", {k: round(v / 1e3, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
for n in path:
    i = calls[n]["in"]
    shp = {k: v.get("shape") for k, v in i.items() if isinstance(v, dict) and "shape" in v}
    print(f"  {n:4d} {calls[n]['op']:11s} {dur[n]:8.1f} us  {shp} {({k: i[k] for k in ('K', 'npoint', 'nsample') if k in i})}")

# ---- per (op, shape) table of the whole step (same measurements) ----
agg = collections.defaultdict(list)
for n, c in enumerate(calls):
    i = c["in"]
    shp = tuple((k, tuple(v["shape"])) for k, v in i.items() if isinstance(v, dict) and "shape" in v)
    extra = tuple((k, i[k]) for k in ("K", "npoint", "nsample") if k in i)
    agg[(c["op"], shp, extra)].append(dur[n])
print("\nper (op, shape): count, mean us, total us")
for (op, shp, extra), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    s = " ".join(f"{k}{list(sh)}" for k, sh in shp)
    print(f"  {op:11s} x{len(v):3d} {np.mean(v):8.1f} us  {sum(v):9.1f} us  {s} {dict(extra) if extra else ''}")
