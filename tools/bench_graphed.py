#!/usr/bin/env python
"""Time the graph-captured fluid train step (tpugan_b200.graph_step) with / without the one-search IDGCN layer."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
import torch, refstep
out = {}
for fused in (False, True):
    ctx = refstep.build("fluid", B=8, n_lo=2048, ratio=4, backend="cuda", capturable=True)
    gs = refstep.graphed_step(ctx, capture=True, fused_idgcn=fused)
    n = 12
    for _ in range(3):
        n += 2; gs.step(n)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        n += 2; l = gs.step(n)
    b.record(); torch.cuda.synchronize()
    out["fused_idgcn" if fused else "reference_idgcn"] = {"ms_per_step": a.elapsed_time(b) / 10, "losses": l}
    ctx.hook.remove(); del gs, ctx
    torch.cuda.empty_cache()
print(json.dumps(out))
