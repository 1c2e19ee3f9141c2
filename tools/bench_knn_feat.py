#!/usr/bin/env python
"""Phase breakdown of knn_feat_tc_kernel (clock64 stamps written by thread 0 of every CTA).  TPG_KNN_DBG=1."""
import ctypes, os, sys
os.environ["TPG_KNN_DBG"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
from tpugan_b200 import _lib
lib = _lib.load()
for (B, P, D, K) in [(8, 2048, 32, 20), (8, 2048, 64, 12), (8, 2048, 32, 9), (8, 2048, 64, 4)]:
    x = torch.randn(B, P, D, device="cuda")
    n = lib.tpg_knn_workspace_bytes(B, P, P, D, K)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d = torch.empty(B, P, K, device="cuda"); i = torch.empty(B, P, K, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.tpg_knn_f32(x.data_ptr(), x.data_ptr(), None, None, B, P, P, D, K, d.data_ptr(), i.data_ptr(), ws.data_ptr(), n, st)
        b.record(); torch.cuda.synchronize()
        assert rc == 0, lib.tpg_last_error()
    nct = B * ((P + 127) // 128)
    dbg_bytes = ((16 * 8 * nct + 255) // 256) * 256
    dbg = ws[n - dbg_bytes: n - dbg_bytes + 128 * nct].view(torch.int64).view(nct, 16).cpu().numpy()
    ph = np.diff(dbg[:, :6], axis=1)
    names = ["setup", "pass0", "tau0-select", "pass1", "drain"]
    print(f"B={B} P={P} D={D} K={K}: call {a.elapsed_time(b) * 1e3:.1f} us; per-CTA cycles (mean/max): " +
          "  ".join(f"{nm}={ph[:, j].mean():.0f}/{ph[:, j].max():.0f}" for j, nm in enumerate(names)) +
          f"  total={(dbg[:, 5] - dbg[:, 0]).mean():.0f}  span={(dbg[:, 5].max() - dbg[:, 0].min())}")
    print('   admission-bound phase: merge+store / barrier wait / sorting network / rest =', (dbg[:, 11] - dbg[:, 2]).mean().astype(int), (dbg[:, 12] - dbg[:, 11]).mean().astype(int), (dbg[:, 13] - dbg[:, 12]).mean().astype(int), (dbg[:, 3] - dbg[:, 13]).mean().astype(int))
    print('   issuer cycles per CTA: wait-full / wait-tmem-free / issue =', dbg[:, 8:11].mean(0).astype(int))
    fo = lib.tpg_knn_fallback_count_offset(B)
    print('   flagged for exact fallback:', int(ws[fo:fo + 4].view(torch.int32)[0]), 'of', B * P)
