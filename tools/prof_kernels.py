#!/usr/bin/env python
"""One launch (after one warm-up) of every hot kernel at a BASELINE configs[1] shape, for
    ncu --set full --clock-control none --import-source on -k regex:'tpg::' -c 40 -f -o gpurun_out/prof_<tag> python tools/prof_kernels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from tpugan_b200 import functional as F  # noqa: E402

rng = np.random.default_rng(1)
dev = torch.device("cuda")
B = 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def twice(fn):
    fn()
    torch.cuda.synchronize()
    flush.fill_(1)
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()


z = np.load(os.path.join(ROOT, "tests", "golden", "knn_real_features.npz"))
x32 = torch.from_numpy(np.repeat(z["c4_K20_D32"], B, 0)).to(dev)   # real generator activations
x64 = torch.from_numpy(np.repeat(z["c14_K12_D64"], B, 0)).to(dev)
twice(lambda: F.knn(x32, x32, 20))     # K2: feat_split, knn_feat_tc, knn_feat_rank, fallback
twice(lambda: F.knn(x64, x64, 12))
p2 = torch.from_numpy(synth.fluid_cloud(rng, B, 2048)).to(dev)
p8 = torch.from_numpy(synth.fluid_cloud(rng, B, 8192)).to(dev)
twice(lambda: F.knn(p2, p2, 20))       # grid kNN (3-D)
twice(lambda: F.frnn(p8, p8, 16, 0.035))
twice(lambda: F.fps(p8, 1024))         # 4-CTA clusters, one-way exchange
twice(lambda: F.fps(p2, 512))
q = p8[:, :1024].contiguous()
twice(lambda: F.ball_query(0.1, 32, p8, q))
f = torch.randn(B, 64, 2048, device=dev)
idx = torch.randint(0, 2048, (B, 2048, 16), device=dev, dtype=torch.int32)
twice(lambda: F.group_fwd(f, idx))
twice(lambda: F.group_reduce_fwd(f, idx, 0))   # K7
go = torch.randn(B, 64, 2048, 16, device=dev)
off, items = F.inverse_index(idx, 2048)
twice(lambda: F.group_bwd(go, off, items, 2048))
twice(lambda: F.inverse_index(idx, 2048))
# K11: flow-embedding conv input; K12: restructured EdgeConv front half
xyz, cen = torch.randn(B, 3, 256, device=dev), torch.randn(B, 3, 256, device=dev)
f2, f1 = torch.randn(B, 256, 256, device=dev), torch.randn(B, 256, 256, device=dev)
idx32 = torch.randint(0, 256, (B, 256, 32), device=dev, dtype=torch.int32)
twice(lambda: F.group_assemble([("gather", xyz, cen), ("gather", f2, None), ("broadcast", f1, None)], idx32))
pq = torch.randn(2, B, 16, 2048, device=dev)
idx20 = torch.randint(0, 2048, (B, 2048, 20), device=dev, dtype=torch.int32)
twice(lambda: F.edge_affine_fwd(pq[0], pq[1], pq[1] - 0.1, idx20, 0.2))
tgt = torch.from_numpy(synth.fluid_cloud(rng, B, 8192)).to(dev)
src = (tgt + 0.003 * torch.randn_like(tgt)).contiguous()
twice(lambda: F.chamfer_fwd(src, tgt, 3))
print("done")
