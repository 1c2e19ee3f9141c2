#!/usr/bin/env python
"""bench.py — TPU-GAN point-neighbourhood hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path of one GAN train step: the exact sequence of boundary
calls (kNN, FRNN, ball query, FPS, gather, grouping fwd/bwd, Chamfer fwd/bwd) that the
reference's unmodified ``tempo_gan_step`` makes (tests/golden/fluid_step_schedule.json,
recorded by tests/golden/make_schedule.py), replayed on synthetic fluid frames of BASELINE
config 2 (2048 -> 8192 particles, batch 8, 3-frame window).  The dense layers between the
calls are not on the path; their outputs are seeded synthetic activations of the recorded
shapes (see tpugan_b200/hotpath_trace.py).

    python bench.py --gpus 1 --steps 10 --warmup 3            # ours, one JSON line
    torchrun ... bench.py --gpus N ...                       # weak scaling: batch 8 per GPU
    python bench.py --impl reference ...                     # CPU oracle arm (host cores)

value   = neighbourhood queries / s with inputs resident in HBM (raw C-ABI path)
e2e     = same metric through the drop-in packages (pytorch3d.ops / frnn / pointnet2_ops /
          chamferdist + autograd) with the step's position frames copied from pinned HOST
          memory each step and the loss read back to the host
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
SITE = os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200")
for _p in (SITE, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "kNN+group queries/s"
UNIT = "queries/s"
GOLDEN = os.path.join(ROOT, "tests", "golden")
# kernel launched by each schedule op (csrc/*.cu) — for the roofline object
OP_KERNEL = {"knn": "knn_feat_tc_kernel", "frnn": "grid_knn_kernel", "ball_query": "ball_query_kernel",
             "fps": "fps_reg_kernel", "gather": "group_fwd_kernel", "group": "group_fwd_kernel",
             "group_bwd": "group_bwd_staged_kernel", "gather_bwd": "group_bwd_kernel", "chamfer": "grid_nn1_kernel",
             "chamfer_bwd": "chamfer_bwd_kernel"}
# GAN-step gradient buckets all-reduced at N > 1 (G / tempo-D / spatial-D parameters)
GRAD_BUCKETS = (439461, 738177, 308737)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fluid", choices=["fluid", "action"])
    ap.add_argument("--batch", type=int, default=None, help="clouds per GPU (default 8 fluid / 8 action)")
    ap.add_argument("--cpu-sample-batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--lanes", type=int, default=64, help="streams of the dependency-aware CUDA-graph leg (1 = off)")
    ap.add_argument("--per-op", action="store_true", help="print the per-op time table to stderr")
    ap.add_argument("--per-call", action="store_true", help="print every call of the schedule with its mean device time")
    return ap.parse_args()


# ------------------------------------------------------------------------------- CPU oracle arm
class OracleOps:
    """Replay back-end over the CPU oracle (NumPy arrays).  Used ONLY for the cpu_baseline
    object and the --impl reference arm; never on the product path."""

    def __init__(self):
        import oracle

        oracle.build()
        self.o = oracle

    def array(self, a):
        return np.ascontiguousarray(a)

    def new_step(self):
        pass

    def tick(self):
        return time.perf_counter()

    def knn(self, p1, p2, K):
        return self.o.knn(p1, p2, K)[1]

    def frnn(self, p1, p2, K, r):
        return self.o.frnn(p1, p2, K, r)[1]

    def fill_negative(self, f, k):
        return np.where(f == -1, k, f)

    def to_i32(self, idx):
        return idx.astype(np.int32)

    def stride_last(self, idx, d):
        return np.ascontiguousarray(idx[:, :, ::d])

    def transpose12(self, x):
        return np.ascontiguousarray(np.swapaxes(x, 1, 2))

    def fps(self, xyz, npoint):
        return self.o.fps(xyz, npoint)

    def gather(self, f, idx, will_bwd=False):
        return np.ascontiguousarray(self.o.group_fwd(f, idx[:, :, None])[..., 0])

    def ball_query(self, r, ns, xyz, new_xyz):
        return self.o.ball_query(r, ns, xyz, new_xyz)

    def group(self, f, idx, will_bwd=False):
        return self.o.group_fwd(f, idx)

    def group_bwd(self, grad_out, idx, N):
        if idx.ndim == 2:
            idx = idx[:, :, None]
        B, C = grad_out.shape[:2]
        return self.o.group_bwd(grad_out.reshape(B, C, idx.shape[1], idx.shape[2]), idx, N)

    def chamfer(self, src, tgt, directions):
        return (src, tgt, directions, self.o.chamfer_fwd(src, tgt, directions))

    def chamfer_bwd(self, h, g):
        src, tgt, directions, r = h
        return self.o.chamfer_bwd(src, tgt, r["i_src"], r["i_tgt"], g, g, directions)[1]

    def finish(self, chamfer, results):
        r = chamfer[3]
        return float((r["sum_src"] + r["sum_tgt"]).mean())


def cpu_replay(workload, batch, steps, warmup):
    """Time `steps` passes of the schedule at `batch` clouds on the host cores (oracle, OpenMP)."""
    from tpugan_b200 import hotpath_trace as ht

    ops = OracleOps()
    ops.o.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    doc = ht.load_schedule(os.path.join(GOLDEN, f"{workload}_step_schedule.json"), batch)
    rp = ht.TraceReplay(doc, ops, seed=1)
    for _ in range(warmup):
        rp.run_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        rp.run_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(queries=rp.queries, s_per_step=dt, cores=ops.o.num_threads(), doc=doc)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch_full = args.batch or 8
    sb = args.cpu_sample_batch
    r = cpu_replay(args.workload, sb, args.steps, args.warmup)
    v = r["queries"] / r["s_per_step"]
    sample = (f"{args.workload} GAN-step schedule at batch {sb} of {batch_full} clouds per step "
              f"({r['queries']} queries/step), C oracle with OpenMP on {r['cores']} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, r["doc"], batch_full, sample_batch=sb),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path = C restatement of the un-vendored native ops (oracle/tpg_oracle.c); the "
                "reference's own extensions are not installable here (SURVEY.md §8c)",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, doc, batch, sample_batch=None):
    c = {
        "workload": ("fluid G+D train-step hot path, 2048->8192 particles, 3-frame window (BASELINE configs[1])"
                     if args.workload == "fluid" else
                     "MSR-action G+D train-step hot path, 128->2048 points, 3-frame window (BASELINE configs[4] shape)"),
        "schedule": f"tests/golden/{args.workload}_step_schedule.json", "batch_per_gpu": batch,
        "n_lo": doc["n_lo"], "n_hi": doc["n_hi"], "calls_per_step": len(doc["calls"]), "op_counts": doc["counts"],
        "parallelism": f"batch-sharded x{args.gpus}",
        "l2": "L2 flushed (256 MiB write) before every timed step",
    }
    if sample_batch is not None:
        c["cpu_sample_batch"] = sample_batch
    return c


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz, self.err = index, False, [], set(), None, None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.err = repr(e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "error": self.err}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------- ours
def run_ours(args):
    # the DAG replay keeps many independent streams busy: use every hardware work queue (must precede CUDA init)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import torch.distributed as dist

    import tpugan_b200
    from tpugan_b200 import _lib, functional as Fn, hotpath_trace as ht

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch or 8
    doc = ht.load_schedule(os.path.join(GOLDEN, f"{args.workload}_step_schedule.json"), batch)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # one flat fp32 buffer = [loss | G grads | tempo-D grads | spatial-D grads]: a single all-reduce per step
    flat = torch.zeros(1 + sum(GRAD_BUCKETS), dtype=torch.float32, device=dev) if world > 1 else None

    def flush_l2():
        flush_buf.fill_(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_step(loss):
        # the GAN step's only exchanges under batch sharding: loss + gradient buckets (NCCL over NVLink)
        if world > 1:
            flat[0:1].copy_(loss.reshape(1))
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            return flat[0].clone()
        return loss

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / max(steps, 1)

    # ---- leg 1: device-resident, raw C-ABI ------------------------------------------------
    ops = ht.TorchCudaOps(dev)
    rp = ht.TraceReplay(doc, ops, seed=1 + rank)
    sampler = ClockSampler(local)

    def step_resident():
        return reduce_step(rp.run_step())

    for _ in range(args.warmup):
        step_resident()
    rp.timers = {}
    sampler.start()
    l0 = tpugan_b200.launch_count()
    ms_res = timed(step_resident, args.steps, 0)
    launches = tpugan_b200.launch_count() - l0
    timers, rp.timers = rp.timers, None
    total_queries = rp.queries * world

    # same step captured once into a CUDA graph and replayed: removes the ~300 host launches per step
    graph_info = None
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                rp.run_step()  # warm the allocator pool on the capture stream
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                graph_loss = rp.run_step()

            def step_graph():
                g.replay()
                return reduce_step(graph_loss)

            ms_graph = timed(step_graph, args.steps, args.warmup)
            graph_info = {"value": total_queries / (ms_graph * 1e-3), "unit": UNIT, "ms_per_step": ms_graph,
                          "note": "identical kernel sequence, one cudaGraphLaunch per step"}
        except Exception as e:  # capture is an optimisation of the host side only
            graph_info = {"error": repr(e)[:300]}
            torch.cuda.synchronize()
    # the same calls again, issued over several streams along the data flow recorded from the reference's step
    # (schedule "deps"): independent chains -- the frames of a window, G vs D passes -- overlap inside one graph
    graph_lanes = None
    if not args.no_graph and args.lanes > 1:
        try:
            torch.cuda.synchronize()
            g2 = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            # overlapped calls: an FPS call should hold 8 SMs, not a 64-SM cluster, for its whole duration
            _lib.set_option("fps.sms_per_cloud", 1)
            _lib.set_option("fps.exclusive_sm", 1)  # ... and keeps those SMs to itself (latency-bound rounds)
            Fn.csr_cache.prefetch_enabled = True     # inverse indices are built off the backward's critical path
            with torch.cuda.stream(side):
                rp.run_step(lanes=args.lanes)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g2):
                lanes_loss = rp.run_step(lanes=args.lanes)
            _lib.set_option("fps.sms_per_cloud", 8)
            _lib.set_option("fps.exclusive_sm", 0)
            Fn.csr_cache.prefetch_enabled = False

            def step_lanes():
                g2.replay()
                return reduce_step(lanes_loss)

            ms_lanes = timed(step_lanes, args.steps, args.warmup)
            graph_lanes = {"value": total_queries / (ms_lanes * 1e-3), "unit": UNIT, "ms_per_step": ms_lanes,
                           "streams": args.lanes, "loss_matches_single_stream": None,
                           "note": "same kernels; issue order relaxed to the recorded data dependencies of the "
                                   "reference's train step, one cudaGraphLaunch per step"}
            if graph_info and "error" not in graph_info:
                g.replay()
                g2.replay()
                torch.cuda.synchronize()
                graph_lanes["loss_matches_single_stream"] = bool(torch.equal(graph_loss, lanes_loss))
        except Exception as e:
            graph_lanes = {"error": repr(e)[:300]}
            _lib.set_option("fps.sms_per_cloud", 8)
            _lib.set_option("fps.exclusive_sm", 0)
            Fn.csr_cache.prefetch_enabled = False
            torch.cuda.synchronize()

    # per-op device time inside the timed region (events recorded around every call)
    op_ms, op_bytes, op_calls = {}, {}, {}
    cursor = {k: 0 for k in timers}
    call_ms = [0.0] * len(doc["calls"])
    for step in range(args.steps):
        for ci, c in enumerate(doc["calls"]):
            op = c["op"]
            a, b = timers[op][cursor[op]]
            cursor[op] += 1
            call_ms[ci] += a.elapsed_time(b)
            op_ms[op] = op_ms.get(op, 0.0) + a.elapsed_time(b)
            op_bytes[op] = op_bytes.get(op, 0) + ht.algorithmic_bytes(c)
            op_calls[op] = op_calls.get(op, 0) + 1
    # split knn by point dimension (3-D search vs feature-space search are different regimes)
    dom = max(op_ms, key=op_ms.get)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    ach = op_bytes[dom] / (op_ms[dom] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": OP_KERNEL[dom], "op": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
        "frac": ach / peak, "traffic": (traffic.get(OP_KERNEL[dom]) or {}).get("dram_bytes_per_launch"),
        "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of one launch of this kernel)",
        "note": ("FPS is a chain of npoint dependent arg-max rounds: latency-bound, its algorithmic bytes are tiny; "
                 "see roofline_hbm_op for the largest bandwidth-bound op" if dom == "fps" else None),
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
        "launches_per_step": op_calls[dom] // max(args.steps, 1),
        "avg_launch_us": op_ms[dom] * 1e3 / op_calls[dom],
        "alg_bytes_per_launch": op_bytes[dom] / op_calls[dom],
        "share_of_step": op_ms[dom] / sum(op_ms.values()),
        "per_op_ms_per_step": {k: v / args.steps for k, v in sorted(op_ms.items(), key=lambda kv: -kv[1])},
        "per_op_gbs": {k: op_bytes[k] / (op_ms[k] * 1e-3) / 1e9 for k in op_ms},
    }
    hbm_ops = [k for k in ("group", "group_bwd", "gather", "gather_bwd") if k in op_ms]
    hop = max(hbm_ops, key=lambda k: op_bytes[k]) if hbm_ops else None
    roofline_hbm = None
    if hop:
        a2 = op_bytes[hop] / (op_ms[hop] * 1e-3) / 1e9
        roofline_hbm = {"bound": "hbm", "kernel": OP_KERNEL[hop], "op": hop, "achieved": a2, "peak": peak, "unit": "GB/s",
                        "frac": a2 / peak, "launches_per_step": op_calls[hop] // max(args.steps, 1),
                        "avg_launch_us": op_ms[hop] * 1e3 / op_calls[hop],
                        "alg_bytes_per_launch": op_bytes[hop] / op_calls[hop],
                        "note": "average over every call of the step incl. small launch-bound ones; per-shape numbers: profiles/*_ops_sweep.md"}
    # the kernel whose removal shortens the DAG-replayed step most (tools/ablate_graph.py): feature-space kNN on
    # tcgen05 -- tensor-bound in its contraction, so report it against the measured dense bf16 peak
    # (tf32 runs at half the bf16 rate)
    roofline_tc = None
    tc_calls = [(ci, c) for ci, c in enumerate(doc["calls"])
                if c["op"] == "knn" and c["in"]["p1"]["shape"][2] in (32, 64) and c["in"]["p2"]["shape"][1] >= 1024
                and int(c["in"]["K"]) <= 24]
    if tc_calls:
        fl = sum(2.0 * c["in"]["p1"]["shape"][0] * c["in"]["p1"]["shape"][1] * c["in"]["p2"]["shape"][1] *
                 c["in"]["p1"]["shape"][2] for _, c in tc_calls)  # algorithmic: one pass over the distance matrix
        ms_tc = sum(call_ms[ci] for ci, _ in tc_calls) / args.steps
        pk = float(peaks.get("bf16_tflops", 1628.7)) / 2.0
        roofline_tc = {"bound": "tensor", "kernel": "knn_feat_tc_kernel", "op": "knn (D = 32 / 64)",
                       "achieved": fl / (ms_tc * 1e-3) / 1e12, "peak": pk, "unit": "TFLOP/s",
                       "frac": fl / (ms_tc * 1e-3) / 1e12 / pk, "launches_per_step": len(tc_calls),
                       "avg_launch_us": ms_tc * 1e3 / len(tc_calls),
                       "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (tf32 rate)",
                       "note": "call = norms + tcgen05 kernel + exact-fallback launch; the kernel issues the contraction "
                               "twice (two-pass candidate selection) and spends two thirds of its time in the SIMT "
                               "selection / exact re-ranking that make the indices bit-exact (DESIGN.md K2)"}
    if args.per_call and rank == 0:
        agg = {}
        for ci, c in enumerate(doc["calls"]):
            sig = c["op"] + " " + " ".join(f"{k}={tuple(v['shape']) if isinstance(v, dict) else v}" for k, v in c["in"].items()
                                          if k not in ("id", "fwd_id"))
            e = agg.setdefault(sig, [0, 0.0])
            e[0] += 1
            e[1] += call_ms[ci] / args.steps
        for sig, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {ms * 1e3:9.1f} us/step  {n:3d} x {ms * 1e3 / n:8.1f} us  {sig}", file=sys.stderr)
    if args.per_op and rank == 0:
        for k, v in roofline["per_op_ms_per_step"].items():
            print(f"  {k:12s} {v:9.3f} ms/step  {op_calls[k] // args.steps:4d} calls  "
                  f"{roofline['per_op_gbs'][k]:9.1f} GB/s alg", file=sys.stderr)

    # ---- leg 2: end to end through the drop-in packages, host buffers ----------------------
    e2e = None
    if not args.no_e2e:
        sops = ht.ShimApiOps(dev)
        rp2 = ht.TraceReplay(doc, sops, seed=1 + rank)
        host_frames = [f.detach().cpu().pin_memory() for f in rp2.frames]
        h2d = sum(f.numel() * f.element_size() for f in host_frames)
        host_loss = torch.empty((), dtype=torch.float32).pin_memory()

        def step_e2e():
            for dst, src in zip(rp2.frames, host_frames):
                dst.copy_(src, non_blocking=True)
            loss = reduce_step(rp2.run_step().detach().float().reshape(()))
            host_loss.copy_(loss, non_blocking=False)  # device -> host read of the step's result
            return host_loss

        ms_e2e = timed(step_e2e, args.steps, args.warmup)
        e2e_mode, ms_e2e_eager = "eager (one host launch per call)", ms_e2e
        e2e_graph_err = None
        if not args.no_graph:
            # the same API calls (host->device frame copies included) captured once with torch.cuda.graph, the way a
            # user of the drop-in packages would wrap a static-shape train step; replayed once per step
            for lanes in ([args.lanes] if args.lanes > 1 else []) + [1]:
                try:
                    torch.cuda.synchronize()
                    g3 = torch.cuda.CUDAGraph()
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    _lib.set_option("fps.sms_per_cloud", 1 if lanes > 1 else 8)
                    _lib.set_option("fps.exclusive_sm", 1 if lanes > 1 else 0)
                    Fn.csr_cache.prefetch_enabled = lanes > 1
                    with torch.cuda.stream(side):
                        rp2.run_step(lanes=lanes)
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g3):
                        for dst, src in zip(rp2.frames, host_frames):
                            dst.copy_(src, non_blocking=True)
                        e2e_loss = rp2.run_step(lanes=lanes).detach().float().reshape(())
                    _lib.set_option("fps.sms_per_cloud", 8)
                    _lib.set_option("fps.exclusive_sm", 0)
                    Fn.csr_cache.prefetch_enabled = False

                    def step_e2e_graph():
                        g3.replay()
                        host_loss.copy_(reduce_step(e2e_loss), non_blocking=False)
                        return host_loss

                    eager_loss = float(step_e2e())
                    graph_loss_v = float(step_e2e_graph())
                    if eager_loss != graph_loss_v:
                        raise RuntimeError(f"captured step differs from eager: {graph_loss_v} vs {eager_loss}")
                    ms_g = timed(step_e2e_graph, args.steps, args.warmup)
                    if ms_g < ms_e2e:
                        ms_e2e = ms_g
                        e2e_mode = ("torch.cuda.graph over the API calls, %d streams along the recorded data dependencies"
                                    % lanes) if lanes > 1 else "torch.cuda.graph over the API calls, single stream"
                    break
                except Exception as e:
                    e2e_graph_err = repr(e)[:200]
                    _lib.set_option("fps.sms_per_cloud", 8)
                    _lib.set_option("fps.exclusive_sm", 0)
                    Fn.csr_cache.prefetch_enabled = False
                    torch.cuda.synchronize()
        e2e = {"value": total_queries / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "issue": e2e_mode,
               "eager_ms_per_step": ms_e2e_eager, "graph_error": e2e_graph_err,
               "api": "pytorch3d.ops.knn_points / frnn.frnn_grid_points / pointnet2_ops.pointnet2_utils.* / "
                      "chamferdist.ChamferDistance + autograd"}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_replay(args.workload, args.cpu_sample_batch, 6, 1)
        cpu_baseline = {"value": r["queries"] / r["s_per_step"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "sample": f"6 passes (+1 warm-up) of the {args.workload} schedule at batch {args.cpu_sample_batch} of "
                                  f"{batch} clouds ({r['queries']} queries and {r['s_per_step']:.2f} s per pass), "
                                  f"C oracle with OpenMP"}

    if rank == 0:
        # headline = the step as the product issues it: one CUDA graph over the recorded data-flow DAG when that
        # capture succeeded, else the single-stream graph, else eager per-call launches (always reported too)
        mode, ms_head = "eager (one host launch per call)", ms_res
        if graph_info and "error" not in graph_info:
            mode, ms_head = "cuda graph, single stream", graph_info["ms_per_step"]
        if graph_lanes and "error" not in graph_lanes and graph_lanes["ms_per_step"] < ms_head:
            mode, ms_head = f"cuda graph, {args.lanes} streams along the recorded data dependencies", graph_lanes["ms_per_step"]
        cfg = workload_config(args, doc, batch)
        cfg["issue"] = mode
        line = {
            "metric": METRIC, "value": total_queries / (ms_head * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_head, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "train_step_hot_path_per_s": world * 1e3 / ms_head,
            "eager": {"value": total_queries / (ms_res * 1e-3), "unit": UNIT, "ms_per_step": ms_res,
                      "note": "same calls launched one by one from Python; the per-op / roofline timings below are "
                              "CUDA events around the calls of this leg"},
            "queries_per_step": total_queries,
            "cuda_graph": graph_info, "cuda_graph_streams": graph_lanes,
            "roofline": roofline, "roofline_hbm_op": roofline_hbm, "roofline_tensor_op": roofline_tc, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
