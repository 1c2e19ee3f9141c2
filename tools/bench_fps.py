#!/usr/bin/env python
"""FPS variant sweep on the GPU box (tuning aid): us per round for (N, points/thread, cluster size)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from tpugan_b200 import _lib
lib = _lib.load()
fn = lib.tpg_debug_fps_variant
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
rng = np.random.default_rng(1)
W = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
def heat():
    for _ in range(40): (W @ W)
    torch.cuda.synchronize()
B = 8
print(torch.cuda.get_device_name(0)); import subprocess; print(subprocess.run(['nvidia-smi','--query-gpu=clocks.sm,clocks.max.sm','--format=csv,noheader'],capture_output=True,text=True).stdout)
for N, npoint in [(2048, 512), (4096, 1024), (8192, 1024), (16384, 1024), (32768, 1024)]:
    xyz = torch.from_numpy(synth.fluid_cloud(rng, B, N)).cuda()
    out = torch.empty((B, npoint), dtype=torch.int32, device="cuda")
    ref = None
    for ppt, cl, flags in [(2, 1, 0), (4, 1, 0), (8, 1, 0), (4, 8, 0), (8, 8, 0), (16, 8, 0), (4, 4, 0), (8, 4, 0), (16, 4, 0), (8, 2, 0), (16, 2, 0)]:
        ppc = N if cl == 1 else ((N + cl - 1) // cl + 31) // 32 * 32
        thr = ((ppc + ppt - 1) // ppt + 31) // 32 * 32
        if thr > 1024 or (cl == 1 and N * 12 > 200000):
            continue
        st = torch.cuda.current_stream().cuda_stream
        rc = fn(xyz.data_ptr(), B, N, npoint, out.data_ptr(), ppt, cl, thr, flags, st)
        if rc != 0:
            print(N, ppt, cl, "rc", rc, lib.tpg_last_error()); continue
        torch.cuda.synchronize()
        if ref is None: ref = out.clone()
        okk = bool((ref == out).all())
        heat()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(xyz.data_ptr(), B, N, npoint, out.data_ptr(), ppt, cl, thr, flags, st); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts)[0] * 1e3
        print(f"N={N:6d} npoint={npoint:5d} ppt={ppt:2d} cl={cl} fl={flags} threads={thr:4d}  {t:8.1f} us  {t / npoint * 1e3:7.1f} ns/round  same={okk}")
