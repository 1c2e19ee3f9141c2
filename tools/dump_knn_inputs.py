#!/usr/bin/env python
"""Dump the feature-space kNN inputs (D = 32 / 64) that the reference's generator produces on real
activations (one SRNet forward, B clouds) -> npz, for off-line analysis of the K2 candidate margins."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
import refstep  # noqa: E402
from tpugan_b200.recording import log  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/knn_inputs.npz"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = refstep.build("fluid", B=B, n_lo=2048, ratio=4, backend="cuda")
for n in range(steps):  # optionally a few optimiser steps first (features of a slightly trained net)
    refstep.step(ctx, 12 + n)
log.start(capture=True)
refstep.generator_forward(ctx)
calls = log.stop()
d = {}
for n, c in enumerate(calls):
    if c.op == "knn" and c.inputs["p1"].shape[2] > 3:
        d[f"c{n}_K{c.inputs['K']}_D{c.inputs['p1'].shape[2]}"] = c.inputs["p1"].cpu().numpy()
np.savez_compressed(out, **d)
print({k: v.shape for k, v in d.items()})
