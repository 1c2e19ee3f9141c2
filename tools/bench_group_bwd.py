#!/usr/bin/env python
"""group_bwd variant sweep on the GPU box (tuning aid): plain CSR gather vs the staged kernel, on the shapes of
the fluid step schedule, idx = real kNN lists of a fluid cloud.  L2 flushed before every call."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from tpugan_b200 import _lib, functional as F
lib = _lib.load()
fn = lib.tpg_debug_group_bwd_variant
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
rng = np.random.default_rng(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B = 8
SHAPES = [(256, 256, 32, 256), (128, 256, 32, 1024), (3, 1024, 32, 2048), (32, 2048, 20, 2048), (32, 2048, 10, 2048),
          (32, 2048, 9, 2048), (64, 2048, 12, 2048), (128, 512, 32, 1024), (64, 2048, 4, 2048), (128, 128, 16, 512),
          (64, 2048, 8, 2048), (64, 8192, 16, 8192), (128, 2048, 32, 2048)]
if len(sys.argv) > 1: SHAPES = [SHAPES[int(sys.argv[1])]]
VARS = (0, 21, 31, 22, 32, 24, 34, -1) if len(sys.argv) <= 2 else (int(sys.argv[2]),)
for C, M, k, N in SHAPES:
    p = torch.from_numpy(synth.fluid_cloud(rng, B, N)).cuda()
    q = p[:, torch.randperm(N, device="cuda")[:M]].contiguous()
    idx = F.knn(q, p, k)[1].to(torch.int32).contiguous()
    L = M * k
    go = torch.randn(B, C, M, k, device="cuda")
    off, items = F.inverse_index(idx, N)
    ref = None
    line = f"C={C:4d} M={M:5d} k={k:3d} N={N:5d} ({4e-6 * B * C * L:6.1f} MB):"
    for var in VARS:
        out = torch.full((B, C, N), float("nan"), device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        args = (go.data_ptr(), off.data_ptr(), items.data_ptr(), B, C, N, L, out.data_ptr(), var, st)
        rc = fn(*args)
        if rc != 0:
            line += f"  v{var}: n/a"; continue
        torch.cuda.synchronize()
        if ref is None: ref = out.clone()
        same = bool(torch.equal(ref, out))
        ts = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(*args); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        t = float(np.median(ts))
        line += f"  v{var}: {t:6.1f}us{'' if same else ' MISMATCH'}"
    print(line, flush=True)
