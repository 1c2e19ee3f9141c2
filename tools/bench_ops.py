#!/usr/bin/env python
"""Per-op microbench sweep on the GPU box (BASELINE configs[2] and configs[3]): device time per call
(CUDA events, median of `reps`, L2 flushed before every call), algorithmic GB/s (SURVEY.md §8d formulas)
and the fraction of the measured HBM peak.  Writes a markdown table.

    python tools/bench_ops.py [--quick] [--out gpurun_out/ops_sweep.md]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from tpugan_b200 import functional as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ops_sweep.md"))
args = ap.parse_args()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rng = np.random.default_rng(1)
rows = []


def timeit(fn, reps=None):
    reps = reps or args.reps
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def add(op, shape, us, alg_bytes, note=""):
    gbs = alg_bytes / (us * 1e-6) / 1e9
    rows.append((op, shape, us, alg_bytes / 1e6, gbs, gbs / PEAK, note))
    print(f"{op:18s} {shape:44s} {us:10.1f} us {alg_bytes / 1e6:10.1f} MB {gbs:8.1f} GB/s {100 * gbs / PEAK:5.1f}%  {note}", flush=True)


def cloud(B, N):
    return torch.from_numpy(synth.fluid_cloud(rng, B, N)).to(dev)


def radius_for(k):  # ~2k points inside the ball at SPH spacing 0.025
    return 0.025 * (2 * k * 3 / (4 * np.pi)) ** (1 / 3)


B = 8
NS = [2048, 8192] if args.quick else [2048, 8192, 32768, 65536]
for N in NS:
    p = cloud(B, N)
    for K in (16, 32):
        us = timeit(lambda: F.knn(p, p, K))
        add("knn (D=3)", f"B={B} N={N} K={K}", us, 4 * B * 3 * 2 * N + 12 * B * N * K, f"{B * N / us:.1f} Mquery/s")
        r = radius_for(K)
        us = timeit(lambda: F.frnn(p, p, K, r))
        add("frnn", f"B={B} N={N} K={K} r={r:.3f}", us, 4 * B * 3 * 2 * N + 12 * B * N * K, f"{B * N / us:.1f} Mquery/s")
        M = N // 4
        q = p[:, :M].contiguous()
        us = timeit(lambda: F.ball_query(r, K, p, q))
        add("ball_query", f"B={B} N={N} M={M} ns={K}", us, 12 * B * (N + M) + 4 * B * M * K)
    if N <= 32768:
        npoint = N // 4
        us = timeit(lambda: F.fps(p, npoint), reps=3)
        add("fps", f"B={B} N={N} npoint={npoint}", us, 12 * B * N + 4 * B * npoint, f"{us / npoint * 1e3:.0f} ns/round (latency-bound)")

for N, D, K in ([(2048, 64, 16)] if args.quick else [(2048, 32, 16), (2048, 64, 16), (8192, 64, 16), (8192, 64, 24)]):
    x = torch.randn(B, N, D, device=dev)
    us = timeit(lambda: F.knn(x, x, K))
    add("knn (tcgen05)", f"B={B} N={N} D={D} K={K}", us, 4 * B * D * 2 * N + 12 * B * N * K,
        f"{2.0 * B * N * N * D * 2 / us / 1e6:.1f} TFLOP/s tf32 issued (2 passes)")

GC = [(64, 2048, 16), (256, 2048, 32)] if args.quick else [(64, 2048, 16), (128, 2048, 32), (256, 2048, 32), (64, 8192, 16), (256, 8192, 32),
                                                            (64, 65536, 16), (128, 65536, 32)]
for C, N, k in GC:
    Bg = B if N <= 8192 else 1
    f = torch.randn(Bg, C, N, device=dev)
    idx = torch.randint(0, N, (Bg, N, k), device=dev, dtype=torch.int32)
    us = timeit(lambda: F.group_fwd(f, idx))
    nbytes = 4 * Bg * (C * N + N * k + C * N * k)
    add("group fwd", f"B={Bg} C={C} N=M={N} k={k}", us, nbytes)
    us = timeit(lambda: F.group_reduce_fwd(f, idx, 0))
    add("group+max fwd", f"B={Bg} C={C} N=M={N} k={k}", us, 4 * Bg * (C * N + N * k + 2 * C * N))
    go = torch.randn(Bg, C, N, k, device=dev)
    off, items = F.inverse_index(idx, N)
    us = timeit(lambda: F.group_bwd(go, off, items, N))
    add("group bwd", f"B={Bg} C={C} N=M={N} k={k}", us, nbytes, "CSR prebuilt")
    us = timeit(lambda: F.inverse_index(idx, N))
    add("inverse index", f"B={Bg} N={N} L={N * k}", us, 4 * Bg * (2 * N * k + N))
    del go

for c, n in ([(64, 8192)] if args.quick else [(64, 8192), (256, 8192), (128, 65536)]):
    m = n // 4
    Bt = B if n <= 8192 else 2
    unk, kn = cloud(Bt, n), cloud(Bt, m)
    us = timeit(lambda: F.three_nn(unk, kn))
    add("three_nn", f"B={Bt} n={n} m={m}", us, 12 * Bt * (n + m) + 24 * Bt * n)
    d, i3 = F.three_nn(unk, kn)
    w = torch.rand(Bt, n, 3, device=dev)
    ff = torch.randn(Bt, c, m, device=dev)
    us = timeit(lambda: F.three_interpolate_fwd(ff, i3, w))
    add("three_interpolate", f"B={Bt} c={c} m={m} n={n}", us, 4 * Bt * c * m + 24 * Bt * n + 4 * Bt * c * n)

# configs[3]: Chamfer fwd + bwd, 8192 x 32768, batch 32
Bc, P1, P2 = (8, 8192, 32768) if args.quick else (32, 8192, 32768)
tgt = cloud(Bc, P2)
src = (tgt[:, ::4] + 0.003 * torch.randn(Bc, P1, 3, device=dev)).contiguous()
g = torch.full((Bc,), 1.0 / Bc, device=dev)
us_f = timeit(lambda: F.chamfer_fwd(src, tgt, 3), reps=3)
add("chamfer fwd", f"B={Bc} {P1}x{P2}", us_f, 20 * Bc * (P1 + P2), f"{2.0 * Bc * P1 * P2 / us_f / 1e3:.1f} Gpair/s brute-force equivalent")
r = F.chamfer_fwd(src, tgt, 3)
us_b = timeit(lambda: F.chamfer_bwd(src, tgt, r["i_src"], r["i_tgt"], g, g, 3), reps=3)
add("chamfer bwd", f"B={Bc} {P1}x{P2}", us_b, 32 * Bc * (P1 + P2))

with open(args.out, "w") as fh:
    fh.write(f"# Op sweep (BASELINE configs[2], configs[3]) — device time per call, algorithmic bytes, % of measured HBM peak ({PEAK} GB/s)\n\n")
    fh.write("CUDA events around single calls, median of %d, L2 flushed before every call, clocks as found.\n\n" % args.reps)
    fh.write("| op | shape | us | alg MB | alg GB/s | % HBM peak | note |\n|---|---|---:|---:|---:|---:|---|\n")
    for op, shape, us, mb, gbs, frac, note in rows:
        fh.write(f"| {op} | {shape} | {us:.1f} | {mb:.1f} | {gbs:.1f} | {100 * frac:.1f}% | {note} |\n")
print("wrote", args.out)
