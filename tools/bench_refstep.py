#!/usr/bin/env python
"""Time the reference's unmodified train step on cuda:0 over the drop-in CUDA packages and split
the device time into hot-path kernels (this library) vs everything else (cuDNN / elementwise /
optimiser kernels of the model).

    python tools/bench_refstep.py [--domain fluid] [--batch 8] [--n-lo 2048] [--ratio 4] [--steps 10]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))


def measure(domain="fluid", batch=8, n_lo=2048, ratio=4, steps=10, warmup=3, patched=False, n_iter=12):
    import torch

    import refstep
    import tpugan_b200
    from tpugan_b200.recording import hot_path_ms, log

    ctx = refstep.build(domain, B=batch, n_lo=n_lo, ratio=ratio, backend="cuda")
    if patched:
        tpugan_b200.patch_reference(ctx.mods)
    for _ in range(warmup):
        refstep.step(ctx, n_iter)
    torch.cuda.synchronize()
    l0 = tpugan_b200.launch_count()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        losses = refstep.step(ctx, n_iter)
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    dev_ms = a.elapsed_time(b) / steps
    launches = (tpugan_b200.launch_count() - l0) // steps
    # one more step with an event pair around every boundary call
    log.start(capture=False, timing=True)
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    refstep.step(ctx, n_iter)
    b2.record()
    calls = log.stop()
    torch.cuda.synchronize()
    per_op = hot_path_ms(calls)
    hot = sum(per_op.values())
    return {
        "domain": domain, "batch": batch, "n_lo": n_lo, "n_hi": n_lo * ratio, "patched": patched,
        "train_steps_per_s": 1e3 / dev_ms, "ms_per_step": dev_ms, "wall_ms_per_step": wall * 1e3,
        "hot_path_ms": hot, "timed_step_ms": a2.elapsed_time(b2), "hot_path_share": hot / a2.elapsed_time(b2),
        "hot_path_per_op_ms": dict(sorted(per_op.items(), key=lambda kv: -kv[1])),
        "boundary_calls": len(calls), "library_launches_per_step": int(launches), "losses": losses,
        "params": refstep.param_counts(ctx),
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--domain", default="fluid")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--n-lo", type=int, default=2048)
    ap.add_argument("--ratio", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--patched", action="store_true")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of one step to this file")
    a = ap.parse_args()
    if a.profile:
        import torch
        from torch.profiler import ProfilerActivity, profile

        import refstep

        ctx = refstep.build(a.domain, B=a.batch, n_lo=a.n_lo, ratio=a.ratio, backend="cuda")
        if a.patched:
            import tpugan_b200

            tpugan_b200.patch_reference(ctx.mods)
        for _ in range(3):
            refstep.step(ctx, 12)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            refstep.step(ctx, 12)
            torch.cuda.synchronize()
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
        sys.exit(0)
    print(json.dumps(measure(a.domain, a.batch, a.n_lo, a.ratio, a.steps, a.warmup, a.patched)))
