"""CPU tests of the oracle itself (run everywhere, `-m "not gpu"`).

The oracle is pinned by (1) fixtures produced by RUNNING the reference's Python
(tests/golden/make_golden.py), (2) independent float64 NumPy brute force, (3) scipy cKDTree
neighbour sets, (4) adversarial inputs (duplicates, 999-dummies, radius boundary, K > P2,
the FPS origin-skip shell)."""
import os

import numpy as np
import pytest

import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


# ------------------------------------------------------------------ golden: reference run live
def test_golden_sampling_py_fps(oracle):
    g = load("fps_sampling_py.npz")
    sel = g["rowsel"]
    for t in "abc":
        idx, rows = oracle.fps_start(g[f"pts_{t}"][None], len(g[f"idx_{t}"]), [int(g[f"start_{t}"])], return_rows=True)
        assert np.array_equal(idx[0], g[f"idx_{t}"]), t
        assert np.array_equal(rows[0][sel], g[f"rows_{t}"]), t  # bit-exact distance rows


def test_golden_index_points(oracle):
    g = load("index_points.npz")
    assert np.array_equal(oracle.gather_rows(g["x"], g["idx2"]), g["out2"])
    B, S, K = g["idx3"].shape
    out3 = oracle.gather_rows(g["x"], g["idx3"].reshape(B, S * K)).reshape(B, S, K, -1)
    assert np.array_equal(out3, g["out3"])


def test_golden_interp_kernels(oracle):
    g = load("interp_kernels.npz")
    # torch's CPU sqrt is Sleef's vectorised sqrt_u05 (not correctly rounded: 1 ulp off in ~0.6% of
    # cases, measured); the oracle uses IEEE sqrtf like torch's CUDA kernel, so allow exactly 1 ulp.
    l2 = oracle.l2dist(g["a"], g["b"])
    assert np.testing.assert_array_max_ulp(l2, g["l2"][:, 0], maxulp=1) is not None
    assert (l2 == g["l2"][:, 0]).mean() > 0.98
    assert (l2[:50] == 0).all()  # coincident points: clamp branch (interpolation.py:13)
    w = oracle.bicubic(g["r"], float(g["cutoff"]))
    np.testing.assert_allclose(w, g["w"][:, 0], rtol=2e-7, atol=0)  # torch.pow(q,3) vs (q*q)*q: <= 1 ulp


# ------------------------------------------------------------------ golden: reference over shims
@pytest.mark.parametrize("tag", ["dense", "pad", "sparse"])
def test_golden_cubic_interpolation(oracle, tag):
    g = load("cubic_interp.npz")
    out = oracle.cubic_interp(g[f"q_{tag}"][None], g[f"field_{tag}"][None], g[f"pos_{tag}"][None],
                              float(g[f"cutoff_{tag}"]))[0]
    np.testing.assert_allclose(out, g[f"out_{tag}"], rtol=1e-5, atol=1e-6)


def test_golden_ball_query_wrapper_is_knn(oracle):
    """discriminator.py:24-40: FRNN(K) with -1 filled from kNN(K) at the same slot == kNN(K)."""
    g = load("ball_query_wrapper.npz")
    K = int(g["sample"])
    _, knn_idx = oracle.knn(g["xyz1"], g["xyz2"], K)
    assert np.array_equal(knn_idx, g["idx"])
    _, fr = oracle.frnn(g["xyz1"], g["xyz2"], K, float(g["radius"]))
    assert (fr == -1).any(), "fixture must exercise the fill"
    assert np.array_equal(np.where(fr == -1, knn_idx, fr), g["idx"])


def test_golden_masking_loss_and_sr_loss(oracle):
    g = load("masking_loss.npz")
    gt, lo, mask = g["gt"], g["lo"], g["mask"]
    _, nbr = oracle.frnn(lo, gt, 1, 0.025 * 1.9)
    _, selfn = oracle.frnn(gt, gt, 16, 0.025 * 1.4)
    cnt = ((selfn != -1).sum(-1) > 3).astype(np.float32)
    cnt = np.concatenate([cnt, np.zeros((cnt.shape[0], 1), np.float32)], 1)
    picked = oracle.gather_rows(cnt[..., None], nbr[..., 0])  # -1 -> appended zero row
    ml = np.abs(mask - picked).mean(dtype=np.float64)
    assert (nbr == -1).any()
    np.testing.assert_allclose(ml, g["masking_loss"], rtol=1e-5)
    cd = oracle.chamfer_distance(gt, g["pred"], bidirectional=True)
    np.testing.assert_allclose(cd, g["cd"], rtol=1e-5)
    np.testing.assert_allclose(cd + 100.0 * ml, g["total"], rtol=1e-5)
    r = oracle.chamfer_fwd(gt, g["pred"], 3)
    gB = np.full((gt.shape[0],), 1.0 / gt.shape[0], np.float32)
    _, gp = oracle.chamfer_bwd(gt, g["pred"], r["i_src"], r["i_tgt"], gB, gB, 3)
    np.testing.assert_allclose(gp, g["grad_pred"], rtol=1e-5, atol=1e-7)


def test_golden_dilated_knn_and_group_max(oracle):
    g = load("dilated_knn.npz")
    pts = np.ascontiguousarray(np.swapaxes(g["feat"], 1, 2))
    _, idx = oracle.knn(pts, pts, 10)
    assert np.array_equal(idx[:, :, ::2], g["idx"])
    g2 = load("idgcn_group_max.npz")
    _, i9 = oracle.knn(pts, pts, 9)
    assert np.array_equal(i9, g2["idx9"])
    out, arg = oracle.group_reduce_fwd(g2["feat"], i9.astype(np.int32), 0)
    assert np.array_equal(out[..., None], g2["out"])


# ------------------------------------------------------------------ independent checks
def _brute64(p1, p2):
    d = p1[:, :, None, :].astype(np.float64) - p2[:, None, :, :].astype(np.float64)
    return (d * d).sum(-1)


@pytest.mark.parametrize("D,K", [(3, 16), (32, 9), (64, 12), (2, 5)])
def test_knn_against_float64_bruteforce(oracle, D, K):
    rng = np.random.default_rng(D * 100 + K)
    p1 = rng.standard_normal((2, 150, D)).astype(np.float32)
    p2 = rng.standard_normal((2, 400, D)).astype(np.float32)
    d, i = oracle.knn(p1, p2, K)
    ref = _brute64(p1, p2)
    order = np.argsort(ref, axis=-1, kind="stable")[:, :, :K]
    # tie-free random data: the fp32 ranking may differ from fp64 only where gaps are < fp32 eps
    srt = np.take_along_axis(ref, order, -1)
    safe = np.ones(order.shape, bool)
    gap = np.diff(np.sort(ref, -1)[:, :, :K + 1], axis=-1)
    safe &= gap > 1e-5 * np.maximum(srt, 1e-12)
    safe[:, :, 1:] &= safe[:, :, :-1]
    assert safe.mean() > 0.95
    assert np.array_equal(i[safe], order[safe])
    np.testing.assert_allclose(d, np.take_along_axis(ref, i, -1), rtol=1e-5, atol=1e-9)
    assert (np.diff(d, axis=-1) >= 0).all()


def test_knn_ties_lowest_index_and_padding(oracle):
    p2 = synth.lattice_cloud(1, 5)  # 125 points, massive ties
    d, i = oracle.knn(p2, p2, 7)
    assert np.array_equal(i[0, :, 0], np.arange(125))  # self first (d=0)
    for q in range(125):
        for s in range(6):
            assert (d[0, q, s], i[0, q, s]) < (d[0, q, s + 1], i[0, q, s + 1])  # strict (d2, idx) order
    # duplicates: the lower index wins
    p = np.zeros((1, 6, 3), np.float32)
    d, i = oracle.knn(p, p, 4)
    assert np.array_equal(i[0], np.tile(np.arange(4), (6, 1)))
    # K > P2: pad (0, 0)
    d, i = oracle.knn(p[:, :2], p[:, :3] + 1, 5)
    assert np.array_equal(i[0, 0], [0, 1, 2, 0, 0]) and (d[0, :, 3:] == 0).all()
    # ragged
    rng = np.random.default_rng(3)
    a = rng.standard_normal((2, 20, 3)).astype(np.float32)
    d, i = oracle.knn(a, a, 4, lengths1=[20, 5], lengths2=[20, 3])
    assert (i[1, 5:] == 0).all() and (d[1, 5:] == 0).all()
    assert (i[1, :5, :3] < 3).all() and (i[1, :5, 3] == 0).all()


def test_frnn_against_ckdtree_sets(oracle):
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(5)
    p = synth.fluid_cloud(rng, 1, 2000)
    r, K = 0.035, 64
    d, i = oracle.frnn(p, p, K, r)
    tree = cKDTree(p[0].astype(np.float64))
    checked = 0
    for q in range(0, 2000, 7):
        got = set(i[0, q][i[0, q] >= 0].tolist())
        d64 = ((p[0].astype(np.float64) - p[0, q].astype(np.float64)) ** 2).sum(-1)
        sure_in = set(np.nonzero(d64 < (r * r) * (1 - 1e-5))[0].tolist())
        maybe = set(np.nonzero(d64 < (r * r) * (1 + 1e-5))[0].tolist())
        assert sure_in <= got <= maybe
        assert set(tree.query_ball_point(p[0, q].astype(np.float64), r * (1 - 1e-5))) <= got
        checked += 1
    assert checked > 100
    valid = i >= 0
    assert (d[valid] < np.float32(r) * np.float32(r)).all() and (d[~valid] == -1).all()
    # strict boundary: a point at exactly distance r is excluded
    a = np.zeros((1, 1, 3), np.float32)
    b = np.array([[[0.5, 0, 0], [0.25, 0, 0]]], np.float32)
    _, ii = oracle.frnn(a, b, 2, 0.5)
    assert ii.tolist() == [[[1, -1]]]


def test_ball_query_semantics(oracle):
    rng = np.random.default_rng(9)
    xyz = synth.fluid_cloud(rng, 2, 600)
    new = xyz[:, ::5].copy()
    new[0, 0] += 5.0  # no hit -> all zeros
    r, ns = 0.04, 8
    out = oracle.ball_query(r, ns, xyz, new)
    for b in range(2):
        for m in range(new.shape[1]):
            diff = xyz[b] - new[b, m]
            d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
            hits = np.nonzero(d2 < np.float32(r) * np.float32(r))[0][:ns]
            exp = np.zeros(ns, np.int64) if len(hits) == 0 else np.concatenate(
                [hits, np.full(ns - len(hits), hits[0])])
            assert np.array_equal(out[b, m], exp), (b, m)
    assert (out[0, 0] == 0).all()


def test_fps_pointnet2_quirks(oracle):
    rng = np.random.default_rng(11)
    p = synth.fluid_cloud(rng, 1, 500)
    p[0, 10:40] *= 0.01  # inside the |p|^2 <= 1e-3 shell: never selected
    idx = oracle.fps(p, 100)[0]
    assert idx[0] == 0 and len(set(idx.tolist())) == 100
    assert not (set(range(10, 40)) & set(idx[1:].tolist()))
    # NumPy restatement
    t = np.full(500, 1e10, np.float32)
    live = (p[0, :, 0] ** 2 + p[0, :, 1] ** 2 + p[0, :, 2] ** 2) > np.float32(1e-3)
    cur, exp = 0, [0]
    for _ in range(99):
        diff = p[0] - p[0, cur]
        d = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        t = np.where(live, np.minimum(t, d), t)
        cand = np.where(live, t, -1.0)
        cur = int(np.argmax(cand))
        exp.append(cur)
    assert idx.tolist() == exp


def test_group_and_reduce_and_backward(oracle):
    rng = np.random.default_rng(13)
    B, C, N, M, k = 2, 5, 40, 30, 6
    f = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, M, k)).astype(np.int32)
    out = oracle.group_fwd(f, idx)
    exp = np.stack([f[b][:, idx[b]] for b in range(B)])
    assert np.array_equal(out, exp)
    cen = rng.standard_normal((B, C, M)).astype(np.float32)
    assert np.array_equal(oracle.group_fwd(f, idx, cen), exp - cen[..., None])
    go = rng.standard_normal(out.shape).astype(np.float32)
    g = oracle.group_bwd(go, idx, N)
    ref = np.zeros((B, C, N), np.float64)
    for b in range(B):
        np.add.at(ref[b], (slice(None), idx[b].reshape(-1)), go[b].reshape(C, -1).astype(np.float64))
    np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-6)
    mx, arg = oracle.group_reduce_fwd(f, idx, 0)
    assert np.array_equal(mx, exp.max(-1)) and np.array_equal(arg, exp.argmax(-1))
    sm, _ = oracle.group_reduce_fwd(f, idx, 1)
    np.testing.assert_allclose(sm, exp.sum(-1), rtol=1e-5, atol=1e-6)


def test_three_nn_interpolate(oracle):
    rng = np.random.default_rng(17)
    u = rng.standard_normal((2, 50, 3)).astype(np.float32)
    kn = rng.standard_normal((2, 20, 3)).astype(np.float32)
    d, i = oracle.three_nn(u, kn)
    ref = np.sqrt(_brute64(u, kn))
    order = np.argsort(ref, -1, kind="stable")[:, :, :3]
    assert np.array_equal(i, order)
    np.testing.assert_allclose(d, np.take_along_axis(ref, order, -1), rtol=1e-5)
    f = rng.standard_normal((2, 4, 20)).astype(np.float32)
    w = rng.uniform(size=(2, 50, 3)).astype(np.float32)
    out = oracle.three_interpolate_fwd(f, i, w)
    exp = np.stack([(f[b][:, i[b]] * w[b][None]).sum(-1) for b in range(2)])
    np.testing.assert_allclose(out, exp, rtol=1e-5, atol=1e-6)


def test_chamfer_against_float64(oracle):
    rng = np.random.default_rng(19)
    s = rng.standard_normal((3, 80, 3)).astype(np.float32)
    t = rng.standard_normal((3, 120, 3)).astype(np.float32)
    r = oracle.chamfer_fwd(s, t, 3)
    ref = _brute64(s, t)
    np.testing.assert_allclose(r["sum_src"], ref.min(2).sum(1), rtol=1e-5)
    np.testing.assert_allclose(r["sum_tgt"], ref.min(1).sum(1), rtol=1e-5)
    assert np.array_equal(r["i_src"], ref.argmin(2)) and np.array_equal(r["i_tgt"], ref.argmin(1))
    # size-independent property: chamfer(x, x) == 0 with identity assignment
    r0 = oracle.chamfer_fwd(s, s, 3)
    assert (r0["sum_src"] == 0).all() and np.array_equal(r0["i_src"], np.tile(np.arange(80), (3, 1)))
    # gradient vs finite differences of the float64 value
    g = np.ones((3,), np.float32)
    gs, gt = oracle.chamfer_bwd(s, t, r["i_src"], r["i_tgt"], g, g, 3)

    def val(ss, tt):
        m = _brute64(ss, tt)
        return m.min(2).sum() + m.min(1).sum()

    eps = 1e-3
    for (b, p, c) in [(0, 3, 0), (1, 40, 2), (2, 79, 1)]:
        sp, sm = s.copy(), s.copy()
        sp[b, p, c] += eps
        sm[b, p, c] -= eps
        fd = (val(sp, t) - val(sm, t)) / (sp[b, p, c].astype(np.float64) - sm[b, p, c])
        np.testing.assert_allclose(gs[b, p, c], fd, rtol=2e-2, atol=1e-3)


def test_ball_query_wrapper_equals_plain_knn():
    """discriminator.py:24-40 (`ball_query_wrapper`): FRNN(K, r) with its -1 slots filled from kNN(K) is the kNN
    list itself — the in-radius hits are a prefix of it in the canonical (d2, index) order."""
    import oracle
    import synth

    rng = np.random.default_rng(3)
    for B, P1, P2, K, r in [(2, 256, 256, 32, 0.08), (2, 300, 500, 16, 0.05), (1, 100, 20, 32, 0.2), (2, 256, 1024, 32, 0.03)]:
        a, b = synth.fluid_cloud(rng, B, P1), synth.fluid_cloud(rng, B, P2)
        _, fi = oracle.frnn(a, b, K, r)
        _, ki = oracle.knn(a, b, K)
        np.testing.assert_array_equal(np.where(fi == -1, ki, fi), ki)


# ---- independent cross-check: the pure-torch dense formulations select the same neighbours -----------------
# (a different algorithm -- expanded-form distance matrix in float64 + sort / top-k -- on tie-free data; the C
# oracle evaluates the canonical fp32 expression on the same float32 inputs)
def _tf():
    from oracle import torch_formulations as tf

    return tf


@pytest.mark.parametrize("B,P1,P2,D,K", [(2, 300, 400, 3, 16), (1, 257, 513, 32, 9), (1, 128, 1024, 64, 20)])
def test_oracle_knn_equals_torch_square_distance_topk(oracle, B, P1, P2, D, K):
    import torch

    rng = np.random.default_rng(5 + D)
    a = rng.standard_normal((B, P1, D)).astype(np.float32)
    b = rng.standard_normal((B, P2, D)).astype(np.float32)
    od, oi = oracle.knn(a, b, K)
    td, ti = _tf().knn(torch.from_numpy(a).double(), torch.from_numpy(b).double(), K)
    np.testing.assert_array_equal(oi, ti.numpy())
    np.testing.assert_allclose(od, td.numpy(), rtol=1e-5, atol=1e-6)


def test_oracle_ball_query_equals_torch_query_ball_point(oracle):
    import torch

    rng = np.random.default_rng(9)
    xyz = synth.fluid_cloud(rng, 2, 1500)
    new_xyz = np.ascontiguousarray(xyz[:, ::6])  # every centre is a cloud point: at least one hit (itself)
    for r, ns in ((0.05, 16), (0.08, 32), (0.03, 8)):
        o = oracle.ball_query(r, ns, xyz, new_xyz)
        t = _tf().query_ball_point(r, ns, torch.from_numpy(xyz).double(), torch.from_numpy(new_xyz).double())
        np.testing.assert_array_equal(o, t.numpy().astype(np.int32))


def test_oracle_fps_equals_torch_farthest_point_sample(oracle):
    import torch

    rng = np.random.default_rng(11)
    xyz = synth.fluid_cloud(rng, 2, 2000) + np.float32(1.0)  # away from the origin: upstream's |p|^2 <= 1e-3 skip is inert
    t = _tf().farthest_point_sample(torch.from_numpy(xyz).double(), 200, start=0).numpy()
    np.testing.assert_array_equal(oracle.fps(xyz, 200), t.astype(np.int32))
    np.testing.assert_array_equal(oracle.fps_start(xyz, 200, np.zeros(2, np.int64)), t)


def test_oracle_grouping_and_chamfer_equal_torch_formulations(oracle):
    import torch

    rng = np.random.default_rng(13)
    f = rng.standard_normal((2, 8, 300)).astype(np.float32)
    idx = rng.integers(0, 300, size=(2, 50, 7)).astype(np.int32)
    np.testing.assert_array_equal(oracle.group_fwd(f, idx), _tf().grouping(torch.from_numpy(f), torch.from_numpy(idx)).numpy())
    a, b = synth.fluid_cloud(rng, 2, 300), synth.fluid_cloud(rng, 2, 500)
    t = float(_tf().chamfer(torch.from_numpy(a).double(), torch.from_numpy(b).double()))
    assert abs(float(oracle.chamfer_distance(a, b, bidirectional=True)) - t) <= 1e-5 * abs(t)


def test_assembly_oracles_equal_the_torch_compositions_of_the_reference(oracle):
    """K11 / K12 restatements vs the reference's own torch expressions: QueryAndGroup / FlowEmbedding's cat
    (discriminator.py:270-277) and EdgeConv's node_affine + edge_affine on grouped tensors (gcn.py:206-211)."""
    import torch

    rng = np.random.default_rng(3)
    B, N, M, k, C = 2, 50, 20, 6, 5
    idx = rng.integers(0, N, (B, M, k)).astype(np.int32)
    pos2 = rng.standard_normal((B, 3, N)).astype(np.float32)
    pos1 = rng.standard_normal((B, 3, M)).astype(np.float32)
    f2 = rng.standard_normal((B, C, N)).astype(np.float32)
    f1 = rng.standard_normal((B, C, M)).astype(np.float32)
    tidx = torch.from_numpy(idx).long()

    def group(f):  # grouping_operation in plain torch
        return torch.stack([torch.from_numpy(f[b])[:, tidx[b]] for b in range(B)])

    ref = torch.cat([group(pos2) - torch.from_numpy(pos1).view(B, -1, M, 1),
                     torch.cat([group(f2), torch.from_numpy(f1).view(B, -1, M, 1).repeat(1, 1, 1, k)], dim=1)], dim=1)
    out = oracle.group_assemble([("gather", pos2, pos1), ("gather", f2, None), ("broadcast", f1, None)], idx)
    assert np.array_equal(out, ref.numpy())

    # EdgeConv: act(Wn f_j + bn) + act(We (f_j - f_i) + be) in float64 vs the restructured form
    Cin, Co = 7, 4
    f = rng.standard_normal((B, Cin, N))
    idn = rng.integers(0, N, (B, N, k)).astype(np.int32)
    Wn, We = rng.standard_normal((Co, Cin)), rng.standard_normal((Co, Cin))
    bn, be = rng.standard_normal(Co), rng.standard_normal(Co)
    lrelu = lambda v: np.where(v > 0, v, 0.2 * v)
    grouped = np.stack([f[b][:, idn[b]] for b in range(B)])             # [B,Cin,N,k]
    edge = grouped - f[:, :, :, None]
    ref64 = lrelu(np.einsum("oc,bcnk->bonk", Wn, grouped) + bn[None, :, None, None]) + \
        lrelu(np.einsum("oc,bcnk->bonk", We, edge) + be[None, :, None, None])
    p = lrelu(np.einsum("oc,bcn->bon", Wn, f) + bn[None, :, None])
    q = np.einsum("oc,bcn->bon", We, f) + be[None, :, None]
    got = oracle.edge_affine_fwd(p.astype(np.float32), q.astype(np.float32), (q - be[None, :, None]).astype(np.float32), idn, 0.2)
    assert np.abs(got - ref64).max() <= 1e-5 * np.abs(ref64).max()
    g = rng.standard_normal(got.shape).astype(np.float32)
    g2, gc = oracle.edge_affine_bwd(g, q.astype(np.float32), (q - be[None, :, None]).astype(np.float32), idn, 0.2)
    pre = np.stack([q[b][:, idn[b]] for b in range(B)]) - (q - be[None, :, None])[:, :, :, None]
    assert np.allclose(g2, np.where(pre > 0, g, 0.2 * g), rtol=1e-6, atol=1e-7) and np.allclose(gc, -g2.sum(-1), rtol=1e-5, atol=1e-5)
