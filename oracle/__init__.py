"""CPU ORACLE — test infrastructure only.

NumPy front-end of ``oracle/tpg_oracle.c`` (the plain-C restatement of the
reference's neighbourhood primitives).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``tpugan_b200``) never does and fails loudly without its CUDA library.

Parity status: **parity unpinned** against upstream binaries (the reference
vendors none of pytorch3d / frnn / pointnet2_ops / chamferdist / dgl and ships
no tests); pinned instead by the reference's call sites, by live runs of the
reference's importable Python (``tests/golden/make_golden.py``) and by
independent float64 / cKDTree checks (``tests/test_oracle.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtpg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc, seconds)."""
    src = os.path.join(_HERE, "tpg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libtpg_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _ci(*v):
    return [ctypes.c_int(int(x)) for x in v]


# --------------------------------------------------------------------------- kNN / FRNN
def knn(p1, p2, K, lengths1=None, lengths2=None):
    """knn_points semantics (gcn_lib/pointnet/gcn.py:13-22): (dists [B,P1,K] f32, idx int64)."""
    p1, p2 = _f32(p1), _f32(p2)
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    l1 = None if lengths1 is None else _i64(lengths1)
    l2 = None if lengths2 is None else _i64(lengths2)
    d = np.empty((B, P1, K), np.float32)
    i = np.empty((B, P1, K), np.int64)
    lib().orc_knn(_p(p1), _p(p2), _p(l1), _p(l2), *_ci(B, P1, P2, D, K), None, _p(d), _p(i))
    return d, i


def frnn(p1, p2, K, r, lengths1=None, lengths2=None):
    """frnn_grid_points semantics (discriminator.py:27): -1 padded (dists, idx)."""
    p1, p2 = _f32(p1), _f32(p2)
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    rr = np.broadcast_to(np.asarray(r, np.float32), (B,)).astype(np.float32)
    r2 = np.ascontiguousarray(rr * rr, dtype=np.float32)
    l1 = None if lengths1 is None else _i64(lengths1)
    l2 = None if lengths2 is None else _i64(lengths2)
    d = np.empty((B, P1, K), np.float32)
    i = np.empty((B, P1, K), np.int64)
    lib().orc_knn(_p(p1), _p(p2), _p(l1), _p(l2), *_ci(B, P1, P2, D, K), _p(r2), _p(d), _p(i))
    return d, i


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = np.empty((B, M, nsample), np.int32)
    lib().orc_ball_query(_p(xyz), _p(new_xyz), *_ci(B, N, M), ctypes.c_float(radius),
                         ctypes.c_int(nsample), _p(out))
    return out


# --------------------------------------------------------------------------- FPS
def fps(xyz, npoint):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), np.int32)
    lib().orc_fps(_p(xyz), *_ci(B, N, npoint), _p(out))
    return out


def fps_start(pts, k, start, return_rows=False):
    pts = _f32(pts)
    B, N, D = pts.shape
    start = _i64(np.broadcast_to(np.asarray(start, np.int64), (B,)))
    out = np.empty((B, k), np.int64)
    rows = np.empty((B, k, N), np.float32) if return_rows else None
    lib().orc_fps_start(_p(pts), *_ci(B, N, D, k), _p(start), _p(out), _p(rows))
    return (out, rows) if return_rows else out


# --------------------------------------------------------------------------- grouping
def group_fwd(f, idx, center=None):
    f, idx = _f32(f), _i32(idx)
    B, C, N = f.shape
    _, M, k = idx.shape
    c = None if center is None else _f32(center)
    out = np.empty((B, C, M, k), np.float32)
    lib().orc_group_fwd(_p(f), _p(idx), _p(c), *_ci(B, C, N, M, k), _p(out))
    return out


def group_bwd(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M, k = grad_out.shape
    g = np.empty((B, C, N), np.float32)
    lib().orc_group_bwd(_p(grad_out), _p(idx), *_ci(B, C, N, M, k), _p(g))
    return g


def group_reduce_fwd(f, idx, op=0):
    f, idx = _f32(f), _i32(idx)
    B, C, N = f.shape
    _, M, k = idx.shape
    out = np.empty((B, C, M), np.float32)
    arg = np.zeros((B, C, M), np.int32)
    lib().orc_group_reduce_fwd(_p(f), _p(idx), *_ci(B, C, N, M, k, op), _p(out), _p(arg))
    return out, arg


def group_reduce_bwd(grad_out, idx, arg, N, op=0):
    grad_out, idx, arg = _f32(grad_out), _i32(idx), _i32(arg)
    B, C, M = grad_out.shape
    k = idx.shape[2]
    g = np.empty((B, C, N), np.float32)
    lib().orc_group_reduce_bwd(_p(grad_out), _p(idx), _p(arg), *_ci(B, C, N, M, k, op), _p(g))
    return g


# --------------------------------------------------------------------------- conv-input assembly (K11 / K12)
def group_assemble(parts, idx):
    """The unfused composition the reference writes in torch — QueryAndGroup (discriminator.py:190, upstream
    pointnet2_utils.QueryAndGroup.forward) and FlowEmbedding.forward (discriminator.py:270-277):
    parts = [("gather", src [B,C,N], center [B,C,M] or None) | ("broadcast", src [B,C,M], None)], idx [B,M,k]
    -> cat over channels of  src[idx] (- center)  /  src repeated along k.  Plain fp32 element ops."""
    idx = _i32(idx)
    k = idx.shape[2]
    outs = []
    for mode, src, center in parts:
        src = _f32(src)
        if mode == "gather":
            outs.append(group_fwd(src, idx, center))
        else:
            outs.append(np.repeat(src[:, :, :, None], k, axis=3))
    return np.concatenate(outs, axis=1)


def _edge_pre(q, center, idx):
    q, center, idx = _f32(q), _f32(center), _i32(idx)
    B = q.shape[0]
    qj = np.stack([q[b][:, idx[b]] for b in range(B)])            # [B,C,M,k]
    return qj - center[:, :, :, None]                               # one fp32 subtraction per element


def edge_affine_fwd(p, q, center, idx, slope):
    """EdgeConv after the restructure (gcn_lib/pointnet/gcn.py:206-211): p[j] + LeakyReLU(q[j] - center_i), every
    step a single fp32 operation in this order (sub, mul by slope on the negative side, add)."""
    p, idx = _f32(p), _i32(idx)
    pre = _edge_pre(q, center, idx)
    act = np.where(pre > 0, pre, pre * np.float32(slope)).astype(np.float32)
    pj = np.stack([p[b][:, idx[b]] for b in range(p.shape[0])])
    return pj + act


def edge_affine_bwd(grad_out, q, center, idx, slope):
    """-> (g2 = grad_out * LeakyReLU'(q[j] - center_i), grad_center = -sum_j g2 summed sequentially in j)."""
    g = _f32(grad_out)
    pre = _edge_pre(q, center, idx)
    g2 = np.where(pre > 0, g, g * np.float32(slope)).astype(np.float32)
    acc = np.zeros(g2.shape[:3], np.float32)
    for j in range(g2.shape[3]):
        acc = acc + g2[..., j]
    return g2, -acc


# --------------------------------------------------------------------------- three_nn / interpolate
def three_nn(unknown, known):
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d = np.empty((B, n, 3), np.float32)
    i = np.empty((B, n, 3), np.int32)
    lib().orc_three_nn(_p(unknown), _p(known), *_ci(B, n, m), _p(d), _p(i))
    return d, i


def three_interpolate_fwd(f, idx, w):
    f, idx, w = _f32(f), _i32(idx), _f32(w)
    B, c, m = f.shape
    n = idx.shape[1]
    out = np.empty((B, c, n), np.float32)
    lib().orc_three_interpolate_fwd(_p(f), _p(idx), _p(w), *_ci(B, c, m, n), _p(out))
    return out


def three_interpolate_bwd(grad_out, idx, w, m):
    grad_out, idx, w = _f32(grad_out), _i32(idx), _f32(w)
    B, c, n = grad_out.shape
    g = np.empty((B, c, m), np.float32)
    lib().orc_three_interpolate_bwd(_p(grad_out), _p(idx), _p(w), *_ci(B, c, m, n), _p(g))
    return g


# --------------------------------------------------------------------------- Chamfer
def chamfer_fwd(src, tgt, directions=3, lengths_src=None, lengths_tgt=None):
    src, tgt = _f32(src), _f32(tgt)
    B, P1, D = src.shape
    P2 = tgt.shape[1]
    ls = None if lengths_src is None else _i64(lengths_src)
    lt = None if lengths_tgt is None else _i64(lengths_tgt)
    d_s = np.zeros((B, P1), np.float32)
    i_s = np.zeros((B, P1), np.int32)
    d_t = np.zeros((B, P2), np.float32)
    i_t = np.zeros((B, P2), np.int32)
    s_s = np.zeros((B,), np.float32)
    s_t = np.zeros((B,), np.float32)
    lib().orc_chamfer_fwd(_p(src), _p(tgt), _p(ls), _p(lt), *_ci(B, P1, P2, D, directions),
                          _p(d_s), _p(i_s), _p(d_t), _p(i_t), _p(s_s), _p(s_t))
    return dict(d_src=d_s, i_src=i_s, d_tgt=d_t, i_tgt=i_t, sum_src=s_s, sum_tgt=s_t)


def chamfer_bwd(src, tgt, i_src, i_tgt, g_src, g_tgt, directions=3, lengths_src=None,
                lengths_tgt=None):
    src, tgt = _f32(src), _f32(tgt)
    B, P1, D = src.shape
    P2 = tgt.shape[1]
    ls = None if lengths_src is None else _i64(lengths_src)
    lt = None if lengths_tgt is None else _i64(lengths_tgt)
    gs = np.empty_like(src)
    gt = np.empty_like(tgt)
    lib().orc_chamfer_bwd(_p(src), _p(tgt), _p(ls), _p(lt), _p(_i32(i_src)), _p(_i32(i_tgt)),
                          _p(_f32(g_src)), _p(_f32(g_tgt)), *_ci(B, P1, P2, D, directions),
                          _p(gs), _p(gt))
    return gs, gt


def chamfer_distance(src, tgt, bidirectional=False, reverse=False, batch_reduction="mean",
                     point_reduction="sum"):
    """chamferdist.ChamferDistance.forward value (loss.py:176-181)."""
    directions = 3 if bidirectional else (2 if reverse else 1)
    r = chamfer_fwd(src, tgt, directions)
    B, P1, _ = np.shape(src)
    P2 = np.shape(tgt)[1]

    def red(s, P):
        s = s.astype(np.float32)
        if point_reduction == "mean":
            s = s / np.float32(P)
        if batch_reduction == "mean":
            return np.float32(s.sum(dtype=np.float32) / np.float32(B))
        if batch_reduction == "sum":
            return np.float32(s.sum(dtype=np.float32))
        return s

    f, b = red(r["sum_src"], P1), red(r["sum_tgt"], P2)
    if bidirectional:
        return f + b
    return b if reverse else f


# --------------------------------------------------------------------------- cubic interpolation
def cubic_interp(query, field, pos, cutoff):
    """gcn_lib/interpolation.py:103-123, batched: query [S,Q,3], field [S,P,F], pos [S,P,3]."""
    query, field, pos = _f32(query), _f32(field), _f32(pos)
    S, Q, _ = query.shape
    P, F = field.shape[1], field.shape[2]
    assert F <= 16
    out = np.empty((S, Q, F), np.float32)
    lib().orc_cubic_interp(_p(query), _p(field), _p(pos), *_ci(S, Q, P, F), ctypes.c_float(cutoff),
                           _p(out))
    return out


def l2dist(src, dst):
    """gcn_lib/interpolation.py:11-14 element-wise: src, dst [n,3] -> [n]."""
    src, dst = _f32(src), _f32(dst)
    out = np.empty((src.shape[0],), np.float32)
    lib().orc_l2dist(_p(src), _p(dst), ctypes.c_int(src.shape[0]), _p(out))
    return out


def bicubic(r, cutoff):
    """gcn_lib/interpolation.py:92-100 element-wise."""
    r = _f32(r).reshape(-1)
    out = np.empty_like(r)
    lib().orc_bicubic(_p(r), ctypes.c_int(r.shape[0]), ctypes.c_float(cutoff), _p(out))
    return out


def gather_rows(x, idx):
    """x [B,N,U], idx [B,L] (negative wraps) -> [B,L,U]."""
    x, idx = _f32(x), _i64(idx)
    B, N, U = x.shape
    L = idx.shape[1]
    out = np.empty((B, L, U), np.float32)
    lib().orc_gather_rows(_p(x), _p(idx), *_ci(B, N, U, L), _p(out))
    return out
