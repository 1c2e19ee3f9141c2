"""GPU tests of every reference-facing Python surface of the drop-in packages that
the train-step test does not already cover: return conventions, autograd, exception types, and
``tpugan_b200.patch_reference()`` against the unpatched reference code."""
import os
import sys

import numpy as np
import pytest
import torch

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(a, b, rtol=RTOL):
    a, b = a.detach().cpu().double().numpy(), np.asarray(b, np.float64)
    scale = max(float(np.abs(b).max()), 1e-30)
    assert np.allclose(a, b, rtol=rtol, atol=rtol * scale), float(np.abs(a - b).max())


@pytest.fixture(scope="module")
def pkgs():
    import tpugan_b200

    tpugan_b200.activate()
    import chamferdist
    import frnn
    import pytorch3d.ops as p3d
    from pointnet2_ops import pointnet2_utils as pu

    return dict(p3d=p3d, frnn=frnn, pu=pu, chamferdist=chamferdist)


# ---------------------------------------------------------------------------------------- pytorch3d.ops
def test_knn_points_return_nn_and_namedtuple(pkgs, oracle):
    rng = np.random.default_rng(1)
    a, b = synth.fluid_cloud(rng, 2, 300), synth.fluid_cloud(rng, 2, 500)
    r = pkgs["p3d"].knn_points(cu(a), cu(b), K=7, return_nn=True)
    od, oi = oracle.knn(a, b, 7)
    assert np.array_equal(r.idx.cpu().numpy(), oi) and np.array_equal(r.dists.cpu().numpy(), od)
    assert r.knn.shape == (2, 300, 7, 3)
    assert np.array_equal(r.knn.cpu().numpy(), b[np.arange(2)[:, None, None], oi])
    d, i, nn = r  # callers unpack three values (gcn.py:16)
    assert nn is r.knn and i.dtype == torch.int64


def test_knn_gather_forward_backward_and_length_mask(pkgs):
    rng = np.random.default_rng(2)
    x = cu(rng.standard_normal((2, 40, 5)).astype(np.float32)).requires_grad_(True)
    idx = cu(rng.integers(0, 40, size=(2, 30, 6)).astype(np.int64))
    out = pkgs["p3d"].knn_gather(x, idx)
    ref = x.detach()[torch.arange(2, device="cuda")[:, None, None], idx]
    assert torch.equal(out, ref)
    g = cu(rng.standard_normal((2, 30, 6, 5)).astype(np.float32))
    out.backward(g)
    xr = x.detach().clone().requires_grad_(True)
    xr[torch.arange(2, device="cuda")[:, None, None], idx].backward(g)
    close(x.grad, xr.grad.cpu().numpy())
    # neighbours beyond lengths are zeroed (pytorch3d semantics)
    lengths = torch.tensor([3, 6], device="cuda")
    out2 = pkgs["p3d"].knn_gather(x.detach(), idx, lengths)
    assert bool((out2[0, :, 3:] == 0).all()) and torch.equal(out2[1], ref[1])


@pytest.mark.parametrize("D,K,P1,P2", [(3, 5, 200, 300), (32, 9, 100, 1100)])
def test_knn_points_distance_backward(pkgs, D, K, P1, P2):
    rng = np.random.default_rng(3 + D)
    a = cu(rng.standard_normal((2, P1, D)).astype(np.float32)).requires_grad_(True)
    b = cu(rng.standard_normal((2, P2, D)).astype(np.float32)).requires_grad_(True)
    d, idx, _ = pkgs["p3d"].knn_points(a, b, K=K)
    w = cu(rng.standard_normal((2, P1, K)).astype(np.float32))
    (d * w).sum().backward()
    ar, br = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    nb = br[torch.arange(2, device="cuda")[:, None, None], idx]
    (((ar.unsqueeze(2) - nb) ** 2).sum(-1) * w).sum().backward()
    close(a.grad, ar.grad.cpu().numpy())
    close(b.grad, br.grad.cpu().numpy())


def test_knn_points_distance_backward_masks_padded_slots(pkgs):
    rng = np.random.default_rng(5)
    a = cu(rng.standard_normal((2, 20, 3)).astype(np.float32)).requires_grad_(True)
    b = cu(rng.standard_normal((2, 10, 3)).astype(np.float32)).requires_grad_(True)
    l2 = torch.tensor([10, 4], device="cuda")
    d, idx, _ = pkgs["p3d"].knn_points(a, b, lengths2=l2, K=6)  # cloud 1: slots 4,5 are padding (idx 0 / dist 0)
    assert bool((d[1, :, 4:] == 0).all()) and bool((idx[1, :, 4:] == 0).all())
    d.sum().backward()
    ar, br = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    nb = br[torch.arange(2, device="cuda")[:, None, None], idx]
    dd = ((ar.unsqueeze(2) - nb) ** 2).sum(-1)
    mask = torch.ones_like(dd)
    mask[1, :, 4:] = 0
    (dd * mask).sum().backward()
    close(a.grad, ar.grad.cpu().numpy())
    close(b.grad, br.grad.cpu().numpy())


def test_knn_points_errors(pkgs):
    with pytest.raises(ValueError):
        pkgs["p3d"].knn_points(torch.zeros(2, 4, 3, device="cuda"), torch.zeros(2, 4, 2, device="cuda"), K=1)
    with pytest.raises(ValueError):
        pkgs["p3d"].knn_points(torch.zeros(2, 4, 3, device="cuda"), torch.zeros(3, 4, 3, device="cuda"), K=1)
    with pytest.raises(RuntimeError):  # no CPU path
        pkgs["p3d"].knn_points(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), K=1)


# ---------------------------------------------------------------------------------------- frnn
def test_frnn_return_tuple_gather_and_autograd(pkgs, oracle):
    rng = np.random.default_rng(7)
    a, b = synth.fluid_cloud(rng, 2, 400), synth.fluid_cloud(rng, 2, 600)
    ta, tb = cu(a).requires_grad_(True), cu(b).requires_grad_(True)
    d, i, nn, grid = pkgs["frnn"].frnn_grid_points(ta, tb, K=8, r=0.04, return_nn=True)
    od, oi = oracle.frnn(a, b, 8, 0.04)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.detach().cpu().numpy(), od) and grid is None
    valid = oi >= 0
    assert valid.any() and (~valid).any()
    ref_nn = np.where(valid[..., None], b[np.arange(2)[:, None, None], np.maximum(oi, 0)], 0.0)
    assert np.array_equal(nn.detach().cpu().numpy(), ref_nn.astype(np.float32))
    # gradient of the valid distances and of the gathered rows
    w = cu(rng.standard_normal(d.shape).astype(np.float32))
    tv = cu(valid)
    ((d * w)[tv].sum() + (nn * 0.5).sum()).backward()
    ar, br = cu(a).requires_grad_(True), cu(b).requires_grad_(True)
    safe = cu(np.maximum(oi, 0))
    nb = br[torch.arange(2, device="cuda")[:, None, None], safe]
    dd = ((ar.unsqueeze(2) - nb) ** 2).sum(-1)
    ((dd * w)[tv].sum() + (nb * tv.unsqueeze(-1) * 0.5).sum()).backward()
    close(ta.grad, ar.grad.cpu().numpy())
    close(tb.grad, br.grad.cpu().numpy())


def test_frnn_errors(pkgs):
    with pytest.raises(TypeError):
        pkgs["frnn"].frnn_grid_points(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), K=2, r=0.1)
    with pytest.raises(ValueError):
        pkgs["frnn"].frnn_grid_points(torch.zeros(1, 4, 3, device="cuda"), torch.zeros(1, 4, 3, device="cuda"), K=0, r=0.1)


# ---------------------------------------------------------------------------------------- pointnet2_ops
def test_query_and_group_and_group_all(pkgs, oracle):
    pu = pkgs["pu"]
    rng = np.random.default_rng(11)
    xyz = synth.fluid_cloud(rng, 2, 1200)
    feat = rng.standard_normal((2, 6, 1200)).astype(np.float32)
    txyz, tf = cu(xyz).requires_grad_(True), cu(feat).requires_grad_(True)
    fidx = pu.furthest_point_sample(txyz, 100)
    assert fidx.dtype == torch.int32 and not fidx.requires_grad
    new_xyz = pu.gather_operation(txyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    out = pu.QueryAndGroup(0.06, 16, use_xyz=True)(txyz, new_xyz, tf)
    assert out.shape == (2, 9, 100, 16)
    ofi = oracle.fps(xyz, 100)
    onew = np.ascontiguousarray(xyz[np.arange(2)[:, None], ofi])
    obq = oracle.ball_query(0.06, 16, xyz, onew)
    gx = oracle.group_fwd(np.ascontiguousarray(xyz.transpose(0, 2, 1)), obq) - onew.transpose(0, 2, 1)[..., None]
    gf = oracle.group_fwd(feat, obq)
    assert np.array_equal(out.detach().cpu().numpy(), np.concatenate([gx, gf], 1).astype(np.float32))
    out.sum().backward()  # gradient reaches the positions through grouping AND the centres (gather_operation)
    assert txyz.grad is not None and tf.grad is not None and bool(torch.isfinite(txyz.grad).all())
    close(tf.grad, oracle.group_bwd(np.ones_like(gf), obq, 1200))
    ga = pu.GroupAll(use_xyz=True)(txyz.detach(), None, tf.detach())
    assert ga.shape == (2, 9, 1, 1200)
    assert torch.equal(ga[:, :3, 0], txyz.detach().transpose(1, 2)) and torch.equal(ga[:, 3:, 0], tf.detach())
    assert pu.QueryAndGroup(0.06, 16, use_xyz=False)(txyz.detach(), new_xyz.detach(), tf.detach()).shape == (2, 6, 100, 16)
    assert pu.QueryAndGroup(0.06, 16)(txyz.detach(), new_xyz.detach(), None).shape == (2, 3, 100, 16)


def test_three_nn_three_interpolate_aliases_autograd(pkgs, oracle):
    pu = pkgs["pu"]
    rng = np.random.default_rng(12)
    unk, kn = synth.fluid_cloud(rng, 2, 500), synth.fluid_cloud(rng, 2, 120)
    dist, idx = pu.three_nn(cu(unk), cu(kn))
    od, oi = oracle.three_nn(unk, kn)
    assert np.array_equal(idx.cpu().numpy(), oi) and idx.dtype == torch.int32
    close(dist, od)
    w = 1.0 / (dist + 1e-8)
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    f = cu(rng.standard_normal((2, 7, 120)).astype(np.float32)).requires_grad_(True)
    out = pu.three_interpolate(f, idx, w)
    close(out, oracle.three_interpolate_fwd(f.detach().cpu().numpy(), oi, w.cpu().numpy()))
    g = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(cu(g))
    close(f.grad, oracle.three_interpolate_bwd(g, oi, w.cpu().numpy(), 120))


def test_pointnet2_input_errors(pkgs):
    pu = pkgs["pu"]
    x = torch.zeros(2, 3, 16, device="cuda")
    with pytest.raises(RuntimeError):  # non-contiguous (upstream TORCH_CHECK)
        pu.grouping_operation(x.transpose(1, 2), torch.zeros(2, 4, 2, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError):  # int64 idx
        pu.grouping_operation(x, torch.zeros(2, 4, 2, dtype=torch.int64, device="cuda"))
    with pytest.raises(RuntimeError):  # CPU tensor
        pu.furthest_point_sample(torch.zeros(2, 16, 3), 4)


# ---------------------------------------------------------------------------------------- tpugan_b200 helpers
def test_gcn_dense_helpers(oracle):
    from tpugan_b200 import gcn_dense

    rng = np.random.default_rng(13)
    x = rng.standard_normal((2, 32, 1100)).astype(np.float32)  # [B,C,N]
    tx = cu(x)
    ei = gcn_dense.dense_knn(tx.unsqueeze(-1), k=9)
    od, oi = oracle.knn(np.ascontiguousarray(x.transpose(0, 2, 1)), np.ascontiguousarray(x.transpose(0, 2, 1)), 9)
    assert ei.shape == (2, 2, 1100, 9) and np.array_equal(ei[0].cpu().numpy(), oi)
    assert bool((ei[1] == torch.arange(1100, device="cuda").view(1, -1, 1)).all())
    d, i = gcn_dense.knn_query(9, cu(np.ascontiguousarray(x.transpose(0, 2, 1))))
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    f = tx.clone().requires_grad_(True)
    sel = gcn_dense.batched_index_select(f.unsqueeze(-1), ei[0])
    assert np.array_equal(sel.detach().cpu().numpy(), oracle.group_fwd(x, oi.astype(np.int32)))
    g = rng.standard_normal(sel.shape).astype(np.float32)
    sel.backward(cu(g))
    close(f.grad, oracle.group_bwd(g, oi.astype(np.int32), 1100))
    f2 = tx.clone().requires_grad_(True)
    gm = gcn_dense.group_max(f2, ei[0].to(torch.int32).contiguous())
    ref = oracle.group_fwd(x, oi.astype(np.int32)).max(-1, keepdims=True)
    assert np.array_equal(gm.detach().cpu().numpy(), ref)
    gm.sum().backward()
    fr = tx.clone().requires_grad_(True)
    from tpugan_b200 import functional as F
    F.GroupingOperation.apply(fr, ei[0].to(torch.int32).contiguous()).max(-1)[0].sum().backward()
    close(f2.grad, fr.grad.cpu().numpy())


def test_sampling_numpy_in_numpy_out(oracle):
    from tpugan_b200 import sampling

    rng = np.random.default_rng(14)
    pts = synth.fluid_cloud(rng, 1, 3000)[0]
    idx, rows = sampling.farthest_point_sampling(pts, 64, initial_idx=5)
    oi, orows = oracle.fps_start(pts[None], 64, np.array([5], np.int64), return_rows=True)
    assert isinstance(idx, np.ndarray) and idx.dtype == np.int64 and rows.dtype == np.float32 and rows.shape == (64, 3000)
    assert np.array_equal(idx, oi[0]) and np.array_equal(rows, orows[0])
    # skip_initial: the start is replaced by the point farthest from it (sampling.py:99-103)
    idx2, _ = sampling.farthest_point_sampling(pts, 16, initial_idx=5, skip_initial=True)
    far = int(((pts - pts[5]) ** 2).sum(-1).argmax())
    assert idx2[0] == far and np.array_equal(idx2, oracle.fps_start(pts[None], 16, np.array([far], np.int64))[0])
    ti, tr = sampling.farthest_point_sampling(cu(pts), 16, initial_idx=5, return_distances=False)
    assert ti.is_cuda and tr is None and np.array_equal(ti.cpu().numpy(), oi[0][:16])


def test_interpolate_vel_lst_matches_per_sample_oracle(oracle):
    from argparse import Namespace

    from tpugan_b200 import interpolation

    rng = np.random.default_rng(15)
    B, P, Q = 2, 2500, 2100
    gt_pos = [synth.fluid_cloud(rng, B, P) for _ in range(2)]
    gt_vel = [rng.standard_normal((B, P, 3)).astype(np.float32) for _ in range(2)]
    pred = [np.ascontiguousarray(g[:, :Q] + 0.004 * rng.standard_normal((B, Q, 3)).astype(np.float32)) for g in gt_pos]
    opt = Namespace(R=0.03)
    gt_adv, pred_adv = interpolation.interpolate_vel_lst([cu(p) for p in pred], [cu(p) for p in gt_pos],
                                                         [cu(v) for v in gt_vel], opt, 1.0)
    for f in range(2):
        assert torch.equal(gt_adv[f], cu(gt_vel[f]) * 0.025)
        field = gt_vel[f] * np.float32(0.025)
        close(pred_adv[f], oracle.cubic_interp(pred[f], field, gt_pos[f], 1.6 * 0.03))
        one = interpolation.cubic_interpolation(cu(pred[f][1]), cu(field[1]), cu(gt_pos[f][1]), 1.6 * 0.03)
        assert torch.equal(one, pred_adv[f][1])  # the single-sample signature of the reference gives the same values


def test_chamferdist_reduction_modes(pkgs, oracle):
    rng = np.random.default_rng(16)
    a, b = synth.fluid_cloud(rng, 3, 300), synth.fluid_cloud(rng, 3, 450)
    cd = pkgs["chamferdist"].ChamferDistance()
    for kw in (dict(), dict(bidirectional=True), dict(reverse=True), dict(bidirectional=True, point_reduction="mean"),
               dict(batch_reduction="sum"), dict(bidirectional=True, batch_reduction=None)):
        v = cd(cu(a), cu(b), **kw)
        ref = oracle.chamfer_distance(a, b, **kw)
        close(v, ref)
    with pytest.raises(ValueError):
        cd(cu(a), cu(b[:2]))
    with pytest.raises(TypeError):
        cd(a, cu(b))


# ---------------------------------------------------------------------------------------- patch_reference
def test_patch_reference_equals_unpatched(oracle):
    import refstep
    import tpugan_b200
    import verify_calls
    from tpugan_b200.recording import log

    mods = refstep.import_reference("cuda")
    dis = mods["discriminator"]
    gcn = sys.modules["gcn_lib.pointnet.gcn"]
    rng = np.random.default_rng(17)
    # ball_query_wrapper: FRNN + kNN + fill  ==  one kNN search
    a, b = cu(synth.fluid_cloud(rng, 2, 256)), cu(synth.fluid_cloud(rng, 2, 256))
    ref_idx = dis.ball_query_wrapper(0.05, 32, a, b)
    torch.manual_seed(3)
    layer = gcn.IDGCNLayer(128, 128, bn=False, insn=False, residual=True).cuda()
    x = cu(rng.standard_normal((2, 128, 1100, 1)).astype(np.float32))
    x1 = x.clone().requires_grad_(True)
    y1 = layer(x1)
    y1.square().sum().backward()
    g1 = [p.grad.clone() for p in layer.parameters()]
    layer.zero_grad()
    h = tpugan_b200.patch_reference(mods)
    try:
        assert len(h.applied) >= 5, h.applied
        assert torch.equal(dis.ball_query_wrapper(0.05, 32, a, b), ref_idx)
        x2 = x.clone().requires_grad_(True)
        log.start(capture=True)
        y2 = layer(x2)
        calls = log.stop()
        assert torch.equal(y2, y1)  # max over neighbours is exact: same values, no [B,C,N,9] tensor
        cnt = verify_calls.counts(calls)
        assert cnt.get("group_reduce", 0) == 1 and cnt.get("knn", 0) == 1 and cnt.get("group", 0) == 2, cnt  # one search, not three
        verify_calls.check_log(oracle, calls)
        y2.square().sum().backward()
        close(x2.grad, x1.grad.cpu().numpy(), rtol=1e-4)
        for p, g in zip(layer.parameters(), g1):
            close(p.grad, g.cpu().numpy(), rtol=1e-4)
        # the whole train step under the patches: 9 FRNN calls of the flow embeddings disappear, every call verified
        ctx = refstep.build("fluid", B=2, n_lo=256, ratio=4, backend="cuda", mods=mods)
        log.start(capture=True)
        losses = refstep.step(ctx, 12)
        calls = log.stop()
        assert all(np.isfinite(v) for v in losses.values())
        cnt = verify_calls.counts(calls)
        assert cnt["frnn"] == 2 and cnt["group_reduce"] == 6 and cnt["group_reduce_bwd"] == 6, cnt
        # 105 groupings - 6 fused into gather+max; QueryAndGroup (27) and FlowEmbedding (9) assemble theirs in one pass
        assert verify_calls.grouping_equivalents(calls) == 99 and cnt["group_assemble"] == 36, cnt
        assert cnt["knn"] == 42 - 12, cnt  # 2 IDGCN layers x 3 frames x 2 repeated searches saved
        verify_calls.check_log(oracle, calls)
    finally:
        h.unpatch()
    assert dis.ball_query_wrapper is not None and gcn.IDGCNLayer.forward.__name__ == "forward"


def test_flow_embedding_and_edgeconv_patches(oracle):
    """f3 / f2: FlowEmbedding's conv input assembled in one pass (identical), EdgeConv after the algebraic restructure
    (fp32 reordering: 1e-5) — each against the reference's unpatched layer."""
    import refstep
    import tpugan_b200
    import verify_calls
    from tpugan_b200.recording import log

    mods = refstep.import_reference("cuda")
    dis = mods["discriminator"]
    gcn = sys.modules["gcn_lib.pointnet.gcn"]
    rng = np.random.default_rng(23)
    torch.manual_seed(5)
    # fp32 convolutions: with cuDNN's default TF32 the two orders of operation differ by the TF32 rounding itself
    # (W f_j - W f_i rounds relative to |f|, W (f_j - f_i) relative to |f_j - f_i|: ~3e-4 of the output here)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    B, N, C = 2, 256, 16
    flow = dis.FlowEmbedding(C, [C, 32, 32], sn=False).cuda()
    pos1 = cu(synth.fluid_cloud(rng, B, N).transpose(0, 2, 1))
    pos2 = cu(synth.fluid_cloud(rng, B, N).transpose(0, 2, 1))
    f1, f2 = (cu(rng.standard_normal((B, C, N)).astype(np.float32)) for _ in range(2))
    edge = gcn.EdgeConv(32, 32, k=20, dilation=2, aggregate='max', mlp_layer=True, bn=False, insn=False).cuda()
    edge_sn = gcn.EdgeConv(3, 32, k=12, dilation=1, aggregate='max', mlp_layer=False, bn=False, insn=False, sn=True).cuda()
    edge_bn = gcn.EdgeConv(32, 32, k=8, bn=True).cuda()
    x = cu(rng.standard_normal((B, 32, 1100)).astype(np.float32))
    xp = cu(synth.fluid_cloud(rng, B, 1100).transpose(0, 2, 1))

    def run_flow():
        a, b, c, d = (t.clone().requires_grad_(True) for t in (pos1, pos2, f1, f2))
        _, y = flow(a, b, c, d, 0.05)
        y.square().sum().backward()
        return y.detach(), [t.grad.clone() for t in (a, b, c, d)], [p.grad.clone() for p in flow.parameters()]

    def run_edge(layer, inp):
        layer.zero_grad()
        t = inp.clone().requires_grad_(True)
        y = layer(t)
        y.square().sum().backward()
        return y.detach(), t.grad.clone(), [p.grad.clone() for p in layer.parameters()]

    flow.zero_grad()
    y0, gi0, gp0 = run_flow()
    e0 = run_edge(edge, x)
    edge_sn.eval()  # spectral norm: no power-iteration update between the two runs
    s0 = run_edge(edge_sn, xp)
    b0 = run_edge(edge_bn, x)
    h = tpugan_b200.patch_reference(mods, edgeconv=True)
    try:
        assert any("FlowEmbedding" in a for a in h.applied) and any("EdgeConv" in a for a in h.applied), h.applied
        flow.zero_grad()
        log.start(capture=True)
        y1, gi1, gp1 = run_flow()
        calls = log.stop()
        cnt = verify_calls.counts(calls)
        assert cnt.get("group_assemble") == 1 and cnt.get("group", 0) == 0, cnt
        verify_calls.check_log(oracle, calls)
        assert torch.equal(y1, y0)  # same values into the same convolutions
        for a, b in zip(gi1 + gp1, gi0 + gp0):
            close(a, b.cpu().numpy(), rtol=1e-4)
        log.start(capture=True)
        e1 = run_edge(edge, x)
        calls = log.stop()
        cnt = verify_calls.counts(calls)
        assert cnt.get("edge_affine") == 1 and cnt.get("edge_affine_bwd") == 1 and cnt.get("group", 0) == 0, cnt
        verify_calls.check_log(oracle, calls)
        close(e1[0], e0[0].cpu().numpy(), rtol=1e-5)
        close(e1[1], e0[1].cpu().numpy(), rtol=1e-4)
        for a, b in zip(e1[2], e0[2]):
            close(a, b.cpu().numpy(), rtol=1e-4)
        s1 = run_edge(edge_sn, xp)  # spectral-normalised 1x1 convs, search on positions, single-conv mlp
        close(s1[0], s0[0].cpu().numpy(), rtol=1e-5)
        close(s1[1], s0[1].cpu().numpy(), rtol=1e-4)
        log.start(capture=True)
        b1 = run_edge(edge_bn, x)    # BatchNorm inside the affine branches: not restructurable, reference path taken
        cnt = verify_calls.counts(log.stop())
        assert cnt.get("edge_affine", 0) == 0 and cnt.get("group") == 1, cnt
        close(b1[0], b0[0].cpu().numpy(), rtol=1e-5)
    finally:
        h.unpatch()
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert gcn.EdgeConv.forward.__name__ == "forward" and dis.FlowEmbedding.forward.__name__ == "forward"


# ---------------------------------------------------------------------------------------- f4: GPU data pipeline
def test_gpu_data_pipeline_matches_reference_functions():
    """tpugan_b200.data_pipeline vs the reference's own CPU code (train_utils.sample_patch_with_fps with scipy's KD-tree
    and the numba FPS of sampling.py), fed the same random choices."""
    import refstep
    from tpugan_b200 import data_pipeline as dp

    refstep.import_reference("cuda")
    import train_utils  # the reference's module (baseline/_ref)

    rng = np.random.default_rng(21)
    N = 12000
    frames = []
    base = (rng.uniform(0, 0.6, size=(N, 3)) + np.array([3.0, -1.0, 0.5])).astype(np.float32)
    for t in range(3):
        frames.append({"pos": (base + 0.01 * t * rng.standard_normal((N, 3))).astype(np.float32),
                       "vel": rng.standard_normal((N, 3)).astype(np.float32)})
    center, _, _ = train_utils.normalize_point_cloud(frames[1]["pos"].copy())
    np.random.seed(5)
    ref, patch, fps_idx = train_utils.sample_patch_with_fps(center, 1.0, sample_num=4096, return_free_surface_particles=False,
                                                            return_patch_and_fps_idx=True)
    np.random.seed(5)
    seed_idx = int(np.random.choice(N))
    start = int(np.random.randint(4096))
    gp, gf = dp.sample_patch_with_fps(cu(center), 4096, seed_idx=seed_idx, fps_start=start)
    assert np.array_equal(np.sort(gp.cpu().numpy()), np.sort(patch))      # same patch (the KD-tree orders ties freely)
    assert np.array_equal(gp.cpu().numpy(), patch)                          # and, on tie-free data, the same order
    assert np.array_equal(gf.cpu().numpy(), fps_idx) and len(fps_idx) == 512
    win = dp.fluid_window([{k: cu(v) for k, v in f.items()} for f in frames], sample_num=4096, jitter=0.0,
                          seed_idx=seed_idx, fps_start=start)
    # the centroid is a float32 mean of 12 000 coordinates near 3.0: NumPy's and the GPU's summation orders differ by ~5e-6
    assert np.allclose(win["highres_pos"].cpu().numpy(), ref["patch_pos"], atol=3e-5)
    assert np.allclose(win["lowres_pos"].cpu().numpy(), ref["ds_pos"], atol=3e-5)
    assert np.array_equal(np.sort(win["patch_idx"].cpu().numpy()), np.sort(patch))
    assert win["highres_pos_left"].shape == (4096, 3) and win["lowres_pos_right"].shape == (512, 3)
    assert torch.equal(win["lowres_vel"], cu(frames[1]["vel"])[win["fps_idx"]])
    j = dp.fluid_window([{k: cu(v) for k, v in f.items()} for f in frames], sample_num=4096, jitter=0.003,
                        seed_idx=seed_idx, fps_start=start)
    d = (j["lowres_pos"] - win["lowres_pos"]).std().item()
    assert 0.002 < d < 0.004
