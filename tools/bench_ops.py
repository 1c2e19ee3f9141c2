#!/usr/bin/env python
"""Per-op microbench sweep on the GPU box (BASELINE configs[2] and configs[3]): device time per call
(CUDA events, median of `reps`, L2 flushed before every call), algorithmic GB/s (SURVEY.md §8d formulas),
the fraction of the measured HBM peak and, for the search ops, brute-force-equivalent pair evaluations/s
(the secondary ceiling of §8d: a search is compute-bound on the SIMT pipe unless a grid prunes it).

    python tools/bench_ops.py [--quick] [--out gpurun_out/ops_sweep.md]

`bench.py --workload sweep|chamfer` imports `run_sweep` / `run_chamfer` from here.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


class Bench:
    def __init__(self, reps=7, verbose=True):
        import torch

        self.torch = torch
        self.dev = torch.device("cuda")
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.rng = np.random.default_rng(1)
        self.reps = reps
        self.rows = []
        self.peak, self.peak_source = hbm_peak()
        self.verbose = verbose

    def timeit(self, fn, reps=None):
        torch = self.torch
        reps = reps or self.reps
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            self.flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        return float(np.median(ts))

    def add(self, op, shape, us, alg_bytes, note="", pairs=None, queries=None):
        gbs = alg_bytes / (us * 1e-6) / 1e9
        row = {"op": op, "shape": shape, "us": us, "alg_mb": alg_bytes / 1e6, "alg_gbs": gbs, "frac_hbm": gbs / self.peak,
               "note": note}
        if pairs is not None:
            row["gpairs_per_s"] = pairs / (us * 1e-6) / 1e9
        if queries is not None:
            row["mqueries_per_s"] = queries / us
        self.rows.append(row)
        if self.verbose:
            extra = f" {row['gpairs_per_s']:9.1f} Gpair/s" if pairs is not None else ""
            print(f"{op:18s} {shape:44s} {us:10.1f} us {alg_bytes / 1e6:10.1f} MB {gbs:8.1f} GB/s "
                  f"{100 * gbs / self.peak:5.1f}%{extra}  {note}", file=sys.stderr, flush=True)

    def cloud(self, B, N):
        import synth

        return self.torch.from_numpy(synth.fluid_cloud(self.rng, B, N)).to(self.dev)


def radius_for(k):  # ~2k points inside the ball at SPH spacing 0.025
    return 0.025 * (2 * k * 3 / (4 * np.pi)) ** (1 / 3)


def run_sweep(quick=False, reps=7, verbose=True):
    """BASELINE configs[2]: knn / FRNN / ball_query / FPS / grouping / three_interpolate at N = 2K-64K,
    k = 16/32, C = 64-256.  Returns the Bench (rows in .rows)."""
    from tpugan_b200 import functional as F

    bn = Bench(reps, verbose)
    torch, dev = bn.torch, bn.dev
    B = 8
    NS = [2048, 8192] if quick else [2048, 8192, 32768, 65536]
    for N in NS:
        p = bn.cloud(B, N)
        for K in (16, 32):
            us = bn.timeit(lambda: F.knn(p, p, K))
            bn.add("knn (D=3)", f"B={B} N={N} K={K}", us, 4 * B * 3 * 2 * N + 12 * B * N * K, pairs=float(B) * N * N, queries=B * N)
            r = radius_for(K)
            us = bn.timeit(lambda: F.frnn(p, p, K, r))
            bn.add("frnn", f"B={B} N={N} K={K} r={r:.3f}", us, 4 * B * 3 * 2 * N + 12 * B * N * K, pairs=float(B) * N * N, queries=B * N)
            M = N // 4
            q = p[:, :M].contiguous()
            us = bn.timeit(lambda: F.ball_query(r, K, p, q))
            bn.add("ball_query", f"B={B} N={N} M={M} ns={K}", us, 12 * B * (N + M) + 4 * B * M * K, pairs=float(B) * N * M, queries=B * M)
        if N <= 32768:
            npoint = N // 4
            us = bn.timeit(lambda: F.fps(p, npoint), reps=3)
            bn.add("fps", f"B={B} N={N} npoint={npoint}", us, 12 * B * N + 4 * B * npoint,
                   f"{us / npoint * 1e3:.0f} ns/round (latency-bound)")
    for N, D, K in ([(2048, 64, 16)] if quick else [(2048, 32, 16), (2048, 32, 20), (2048, 64, 12), (2048, 64, 16), (8192, 64, 16), (8192, 64, 24)]):
        x = torch.randn(B, N, D, device=dev)
        us = bn.timeit(lambda: F.knn(x, x, K))
        bn.add("knn (tcgen05)", f"B={B} N={N} D={D} K={K}", us, 4 * B * D * 2 * N + 12 * B * N * K,
               f"{2.0 * B * N * N * D / us / 1e6:.1f} TFLOP/s algorithmic", pairs=float(B) * N * N, queries=B * N)
    GC = [(64, 2048, 16), (256, 2048, 32), (64, 8192, 16)] if quick else [(64, 2048, 16), (128, 2048, 32), (256, 2048, 32), (64, 8192, 16),
                                                          (256, 8192, 32), (64, 65536, 16), (128, 65536, 32)]
    for C, N, k in GC:
        Bg = B if N <= 8192 else 1
        f = torch.randn(Bg, C, N, device=dev)
        idx = torch.randint(0, N, (Bg, N, k), device=dev, dtype=torch.int32)
        us = bn.timeit(lambda: F.group_fwd(f, idx))
        nbytes = 4 * Bg * (C * N + N * k + C * N * k)
        bn.add("group fwd", f"B={Bg} C={C} N=M={N} k={k}", us, nbytes)
        us_g = us
        us = bn.timeit(lambda: F.group_fwd(f, idx).max(-1))
        bn.add("group fwd + max", f"B={Bg} C={C} N=M={N} k={k}", us, 4 * Bg * (C * N + N * k + 2 * C * N), "unfused: materialise [B,C,M,k], torch.max")
        us = bn.timeit(lambda: F.group_reduce_fwd(f, idx, 0))
        bn.add("group+max fused", f"B={Bg} C={C} N=M={N} k={k}", us, 4 * Bg * (C * N + N * k + 2 * C * N))
        go = torch.randn(Bg, C, N, k, device=dev)
        us = bn.timeit(lambda: F.group_bwd_auto(go, idx, N))
        bn.add("group bwd", f"B={Bg} C={C} N=M={N} k={k}", us, nbytes, "inverse index cached")
        us = bn.timeit(lambda: F.inverse_index(idx, N))
        bn.add("inverse index", f"B={Bg} N={N} L={N * k}", us, 4 * Bg * (2 * N * k + N))
        del go
    # K11: conv-input assembly — QueryAndGroup (discriminator.py:190) and FlowEmbedding (discriminator.py:270-277)
    # shapes of the fluid step, one pass vs the reference's composition (groupings, "- centre", repeat, torch.cat)
    for (Cf, N, M, k, flow) in ([(128, 1024, 256, 32, False)] if quick else
                                [(3, 2048, 1024, 32, False), (128, 1024, 256, 32, False), (256, 256, 256, 32, True)]):
        xyz = torch.randn(B, 3, N, device=dev)
        cen = torch.randn(B, 3, M, device=dev)
        f2 = torch.randn(B, Cf, N, device=dev)
        f1 = torch.randn(B, Cf, M, device=dev)
        idx = torch.randint(0, N, (B, M, k), device=dev, dtype=torch.int32)
        parts = [("gather", xyz, cen), ("gather", f2, None)] + ([("broadcast", f1, None)] if flow else [])
        ctot = 3 + Cf * (2 if flow else 1)
        nbytes = 4 * B * (3 * N + 3 * M + Cf * N + (Cf * M if flow else 0) + M * k + ctot * M * k)

        def unfused():
            a = F.group_fwd(xyz, idx) - cen.unsqueeze(-1)
            b_ = F.group_fwd(f2, idx)
            if flow:
                b_ = torch.cat([b_, f1.view(B, -1, M, 1).repeat(1, 1, 1, k)], dim=1)
            return torch.cat([a, b_], dim=1)

        us = bn.timeit(unfused)
        name = "FlowEmbedding input" if flow else "QueryAndGroup"
        bn.add(name + " (unfused)", f"B={B} C=3+{Cf}{'x2' if flow else ''} N={N} M={M} k={k}", us, nbytes, "groupings, subtraction, repeat, torch.cat")
        us = bn.timeit(lambda: F.group_assemble(parts, idx))
        bn.add(name + " (K11)", f"B={B} C=3+{Cf}{'x2' if flow else ''} N={N} M={M} k={k}", us, nbytes)
    # K12: the k-expanded front half of EdgeConv (gcn.py:206-211) at the generator's shapes: reference composition
    # (grouping, "- centre", two 1x1 convolutions + LeakyReLU on [B,C,N,k], add) vs per-node convolutions + one kernel
    for (Cin, Co, N, k) in ([(32, 16, 2048, 20)] if quick else [(32, 16, 2048, 20), (32, 16, 2048, 10), (64, 32, 2048, 12)]):
        feat = torch.randn(B, Cin, N, device=dev)
        idx = torch.randint(0, N, (B, N, k), device=dev, dtype=torch.int32)
        node = torch.nn.Sequential(torch.nn.Conv2d(Cin, Co, 1), torch.nn.LeakyReLU(0.2)).to(dev)
        edge = torch.nn.Sequential(torch.nn.Conv2d(Cin, Co, 1), torch.nn.LeakyReLU(0.2)).to(dev)
        nbytes = 4 * B * (Cin * N + N * k + Co * N * k)

        def ref_front():
            with torch.no_grad():
                g = F.group_fwd(feat, idx)
                return node(g) + edge(g - feat.unsqueeze(-1))

        def restructured():
            with torch.no_grad():
                x = feat.unsqueeze(-1)
                p = node(x).squeeze(-1)
                q = edge[0](x).squeeze(-1)
                c = q - edge[0].bias.view(1, -1, 1)
                return F.edge_affine_fwd(p, q, c, idx, 0.2)

        us = bn.timeit(ref_front)
        bn.add("EdgeConv front (reference)", f"B={B} C={Cin}->{Co} N={N} k={k}", us, nbytes, "grouping, - centre, 2 convs + LeakyReLU on [B,C,N,k], add")
        us = bn.timeit(restructured)
        bn.add("EdgeConv front (K12)", f"B={B} C={Cin}->{Co} N={N} k={k}", us, nbytes, "2 per-node convs + edge_affine kernel")
    for c, n in ([(64, 8192)] if quick else [(64, 8192), (256, 8192), (128, 65536)]):
        m = n // 4
        Bt = B if n <= 8192 else 2
        unk = bn.cloud(Bt, n)
        kn = unk[:, ::4].contiguous()  # feature propagation: the known points are a subset of the unknown cloud
        us = bn.timeit(lambda: F.three_nn(unk, kn))
        bn.add("three_nn", f"B={Bt} n={n} m={m}", us, 12 * Bt * (n + m) + 24 * Bt * n, pairs=float(Bt) * n * m, queries=Bt * n)
        d, i3 = F.three_nn(unk, kn)
        w = torch.rand(Bt, n, 3, device=dev)
        ff = torch.randn(Bt, c, m, device=dev)
        us = bn.timeit(lambda: F.three_interpolate_fwd(ff, i3, w))
        bn.add("three_interpolate", f"B={Bt} c={c} m={m} n={n}", us, 4 * Bt * c * m + 24 * Bt * n + 4 * Bt * c * n)
    return bn


def run_chamfer(quick=False, reps=5, verbose=True, bn=None):
    """BASELINE configs[3]: Chamfer fwd + bwd, 8192 x 32768 points, batch 32 (src = subset of tgt + jitter)."""
    from tpugan_b200 import functional as F

    bn = bn or Bench(reps, verbose)
    torch, dev = bn.torch, bn.dev
    Bc, P1, P2 = (8, 8192, 32768) if quick else (32, 8192, 32768)
    tgt = bn.cloud(Bc, P2)
    src = (tgt[:, ::4] + 0.003 * torch.randn(Bc, P1, 3, device=dev)).contiguous()
    g = torch.full((Bc,), 1.0 / Bc, device=dev)
    us_f = bn.timeit(lambda: F.chamfer_fwd(src, tgt, 3), reps=reps)
    bn.add("chamfer fwd", f"B={Bc} {P1}x{P2}", us_f, 20 * Bc * (P1 + P2), pairs=2.0 * Bc * P1 * P2, queries=Bc * (P1 + P2))
    r = F.chamfer_fwd(src, tgt, 3)
    # gradient w.r.t. src (the prediction); the target cloud has no gradient (loss.py:176-181)
    us_b = bn.timeit(lambda: F.chamfer_bwd(src, tgt, r["i_src"], r["i_tgt"], g, g, 3, need_src=True, need_tgt=False), reps=reps)
    bn.add("chamfer bwd (src)", f"B={Bc} {P1}x{P2}", us_b, 32 * Bc * (P1 + P2))
    us_b2 = bn.timeit(lambda: F.chamfer_bwd(src, tgt, r["i_src"], r["i_tgt"], g, g, 3), reps=reps)
    bn.add("chamfer bwd (both)", f"B={Bc} {P1}x{P2}", us_b2, 32 * Bc * (P1 + P2) + 12 * Bc * P2)
    return bn, dict(B=Bc, P1=P1, P2=P2, us_fwd=us_f, us_bwd=us_b)


def write_markdown(bn, path, reps):
    with open(path, "w") as fh:
        fh.write(f"# Op sweep (BASELINE configs[2], configs[3]) — device time per call, algorithmic bytes, % of measured HBM peak ({bn.peak} GB/s)\n\n")
        fh.write("CUDA events around single calls, median of %d, L2 flushed before every call, clocks as found.\n\n" % reps)
        fh.write("| op | shape | us | alg MB | alg GB/s | % HBM peak | Gpair/s | note |\n|---|---|---:|---:|---:|---:|---:|---|\n")
        for r in bn.rows:
            gp = f"{r['gpairs_per_s']:.1f}" if "gpairs_per_s" in r else ""
            fh.write(f"| {r['op']} | {r['shape']} | {r['us']:.1f} | {r['alg_mb']:.1f} | {r['alg_gbs']:.1f} | {100 * r['frac_hbm']:.1f}% | {gp} | {r['note']} |\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ops_sweep.md"))
    args = ap.parse_args()
    bn = run_sweep(args.quick, args.reps)
    run_chamfer(args.quick, 3, True, bn)
    write_markdown(bn, args.out, args.reps)
    print("wrote", args.out)
