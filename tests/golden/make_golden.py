"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN PYTHON.

Build container only (needs /root/reference; it cannot travel to the GPU box, the fixtures
do).  What the reference can execute by itself pins the oracle directly:

  fps_sampling_py.npz    sampling.farthest_point_sampling (sampling.py:50-106, numba loop
                         :36-44) run LIVE — indices + distance rows
  index_points.npz       discriminator.index_points / loss.index_points (:43-60 / :10-27)
                         incl. the -1 index that wraps to the appended zero row (loss.py:270-275)
  interp_kernels.npz     gcn_lib.interpolation.l2dist (:11-14) and bicubic_kernel (:92-100)

What needs the un-vendored native packages runs over the oracle-backed shims
(oracle/shims): the reference's Python around the boundary is real, the neighbour search
below it is the oracle's — these fixtures pin the *composition* (call-site semantics):

  cubic_interp.npz       gcn_lib.cubic_interpolation (:103-123) incl. the knn-padding branch
  ball_query_wrapper.npz discriminator.ball_query_wrapper (:24-40) == FRNN + kNN fill
  masking_loss.npz       loss.masking_loss (:253-275) and loss.tpugan_sr_loss (:168-183)
  dilated_knn.npz        gcn_lib.pointnet.gcn.DilatedKnnGraph (:74-93) k=20, dilation 2
  idgcn_group_max.npz    grouping_operation + max of IDGCNLayer (gcn.py:258-263)

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle.shims as sh  # noqa: E402

sh.activate()
sys.path.insert(1, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200", "import_stubs"))
sys.path.insert(1, "/root/reference")
warnings.simplefilter("ignore")
torch.Tensor.cuda = lambda self, *a, **k: self  # loss.py:174

import synth  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote %-24s %7d bytes  %s" % (name, os.path.getsize(path), {k: v.shape for k, v in arrays.items()}))


def main():
    rng = np.random.default_rng(20261018)
    torch.manual_seed(1)

    # ---- live: sampling.py --------------------------------------------------------------
    import sampling

    pts = synth.fluid_cloud(rng, 1, 1500)[0]
    idx_a, rows_a = sampling.farthest_point_sampling(pts, 96, initial_idx=7)
    pts2 = synth.with_duplicates(rng, synth.action_cloud(rng, 1, 600))[0]
    idx_b, rows_b = sampling.farthest_point_sampling(pts2, 64, initial_idx=0)
    pts3 = rng.uniform(-1, 1, size=(300, 2)).astype(np.float32)
    idx_c, rows_c = sampling.farthest_point_sampling(pts3, 300, initial_idx=299)
    sel = np.array([0, 1, 17, 63])  # a few full distance rows keep the fixture small
    save("fps_sampling_py.npz", pts_a=pts, start_a=np.int64(7), idx_a=idx_a, rowsel=sel,
         rows_a=rows_a[sel].astype(np.float32), pts_b=pts2, start_b=np.int64(0), idx_b=idx_b,
         rows_b=rows_b[sel].astype(np.float32), pts_c=pts3, start_c=np.int64(299), idx_c=idx_c,
         rows_c=rows_c[sel].astype(np.float32))

    # ---- live: index_points ----------------------------------------------------------------
    import discriminator
    import loss as ref_loss

    x = rng.standard_normal((3, 50, 5)).astype(np.float32)
    i2 = rng.integers(-1, 50, size=(3, 40)).astype(np.int64)
    i3 = rng.integers(0, 50, size=(3, 20, 4)).astype(np.int64)
    o2 = discriminator.index_points(torch.from_numpy(x), torch.from_numpy(i2)).numpy()
    o2b = ref_loss.index_points(torch.from_numpy(x), torch.from_numpy(i2)).numpy()
    assert np.array_equal(o2, o2b)
    o3 = discriminator.index_points(torch.from_numpy(x), torch.from_numpy(i3)).numpy()
    save("index_points.npz", x=x, idx2=i2, out2=o2, idx3=i3, out3=o3)

    # ---- live: interpolation kernels ---------------------------------------------------------
    from gcn_lib import interpolation as interp

    a = (rng.uniform(-1, 1, size=(4000, 3))).astype(np.float32)
    b = (a + rng.normal(0, 0.03, size=a.shape)).astype(np.float32)
    b[:50] = a[:50]  # exact coincidences -> clamp branch
    r = interp.l2dist(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    rr = np.concatenate([np.linspace(0, 0.2, 2001, dtype=np.float32), np.float32([0.08, 0.16, 0.160001, 0.0])])
    w = interp.bicubic_kernel(torch.from_numpy(rr)[:, None], 0.16).numpy()
    save("interp_kernels.npz", a=a, b=b, l2=r, r=rr, w=w, cutoff=np.float32(0.16))

    # ---- composition: cubic_interpolation ---------------------------------------------------
    cases = {}
    for tag, (Q, P, cutoff, far) in {"dense": (400, 600, 0.16, False), "pad": (300, 500, 0.05, True),
                                     "sparse": (200, 200, 0.03, False)}.items():
        pos = synth.fluid_cloud(rng, 1, P)[0]
        q = (pos[rng.integers(0, P, size=Q)] + rng.normal(0, 0.01, size=(Q, 3))).astype(np.float32)
        if far:
            q[:5] += 3.0  # queries with no neighbour -> knn-padding branch (interpolation.py:44-60)
        field = rng.standard_normal((P, 3)).astype(np.float32)
        out = interp.cubic_interpolation(torch.from_numpy(q), torch.from_numpy(field), torch.from_numpy(pos),
                                         cutoff).numpy()
        cases.update({f"q_{tag}": q, f"field_{tag}": field, f"pos_{tag}": pos, f"cutoff_{tag}": np.float32(cutoff),
                      f"out_{tag}": out})
    save("cubic_interp.npz", **cases)

    # ---- composition: ball_query_wrapper ------------------------------------------------------
    xyz2 = synth.fluid_cloud(rng, 2, 700)
    xyz1 = xyz2[:, ::3].copy()
    bq = discriminator.ball_query_wrapper(0.03, 16, torch.from_numpy(xyz1), torch.from_numpy(xyz2)).numpy()
    save("ball_query_wrapper.npz", xyz1=xyz1, xyz2=xyz2, radius=np.float32(0.03), sample=np.int64(16), idx=bq)

    # ---- composition: masking_loss / tpugan_sr_loss -------------------------------------------
    gt = synth.fluid_cloud(rng, 2, 1024)
    lo = (gt[:, ::4] + rng.normal(0, 0.003, size=(2, 256, 3))).astype(np.float32)
    lo[:, :7] += 1.0  # inputs without a gt neighbour -> -1 -> appended zero row
    mask = rng.uniform(size=(2, 256, 1)).astype(np.float32)
    ml = ref_loss.masking_loss(torch.from_numpy(gt), torch.from_numpy(lo), torch.from_numpy(mask), 0.025)
    pred = (gt + rng.normal(0, 0.004, size=gt.shape)).astype(np.float32)[:, :768]
    tp = torch.from_numpy(pred).requires_grad_(True)
    total, cdv, mlv = ref_loss.tpugan_sr_loss(100.0, torch.from_numpy(gt), tp, torch.from_numpy(lo),
                                              torch.from_numpy(mask), 0.025, 12)
    total.backward()
    save("masking_loss.npz", gt=gt, lo=lo, mask=mask, masking_loss=np.float32(ml.item()), pred=pred,
         total=np.float32(total.item()), cd=np.float32(cdv.item()), ml=np.float32(mlv.item()),
         grad_pred=tp.grad.numpy())

    # ---- composition: DilatedKnnGraph + IDGCN gather+max ---------------------------------------
    from gcn_lib.pointnet import gcn as ref_gcn
    from pointnet2_ops.pointnet2_utils import grouping_operation

    feat = rng.standard_normal((2, 32, 300)).astype(np.float32)  # [B,C,N]
    dk = ref_gcn.DilatedKnnGraph(k=10, dilation=2)
    didx = dk(torch.from_numpy(feat).transpose(1, 2).contiguous())  # [B,N,k]
    d_np = didx.numpy() if isinstance(didx, torch.Tensor) else didx[0].numpy()
    _, i9 = ref_gcn.knn_query(9, torch.from_numpy(feat).transpose(1, 2).contiguous())
    g = grouping_operation(torch.from_numpy(feat), i9.type(torch.int32).contiguous())
    gm = torch.max(g, dim=-1, keepdim=True)[0].numpy()
    save("dilated_knn.npz", feat=feat, idx=d_np.astype(np.int64))
    save("idgcn_group_max.npz", feat=feat, idx9=i9.numpy(), out=gm)


if __name__ == "__main__":
    main()
