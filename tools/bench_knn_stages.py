#!/usr/bin/env python
"""Cumulative stage times of the tensor-core kNN call (TPG_KNN_STOP=1..4; median of 20 timed calls each, CUDA events)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
    import numpy as np
    import torch
    from tpugan_b200 import _lib

    lib = _lib.load()
    out = []
    for (B, P, D, K) in [(8, 2048, 32, 20), (8, 2048, 64, 12), (8, 2048, 32, 9)]:
        x = torch.randn(B, P, D, device="cuda")
        n = lib.tpg_knn_workspace_bytes(B, P, P, D, K)
        ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d = torch.empty(B, P, K, device="cuda")
        i = torch.empty(B, P, K, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ts = []
        for rep in range(25):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = lib.tpg_knn_f32(x.data_ptr(), x.data_ptr(), None, None, B, P, P, D, K, d.data_ptr(), i.data_ptr(),
                                 ws.data_ptr(), n, st)
            b.record()
            torch.cuda.synchronize()
            assert rc == 0
            ts.append(a.elapsed_time(b) * 1e3)
        out.append(float(np.median(ts[5:])))
    print(" ".join(f"{t:.1f}" for t in out))
else:
    print("stages: 1 split | 2 + tcgen05 | 3 + rank | 4 + fallback;  columns: D=32 K=20, D=64 K=12, D=32 K=9 (us, cumulative)")
    for stop in (1, 2, 3, 4):
        env = dict(os.environ, TPG_KNN_STOP=str(stop))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(stop, r.stdout.strip(), r.stderr.strip()[-300:])
