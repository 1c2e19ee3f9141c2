"""BASELINE.json's full sizes (configs[1..3]) on the GPU: the oracle would take minutes on whole problems, so each
test checks (i) the CUDA result against the oracle on a slice of the queries / channels with the FULL candidate
set (bit-exact indices), and (ii) size-independent properties over the whole output: sortedness, recomputed
distances, self-neighbour, conservation of sums by the scatter-add backward, symmetry of Chamfer."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def F():
    import tpugan_b200.functional as F_

    return F_


def _check_knn_properties(p1, p2, d, i, K, self_query):
    B, P1, _ = p1.shape
    assert d.shape == (B, P1, K) and i.shape == (B, P1, K)
    assert bool((d[:, :, 1:] >= d[:, :, :-1]).all())                      # ascending
    assert int(i.min()) >= 0 and int(i.max()) < p2.shape[1]
    nb = torch.gather(p2, 1, i.reshape(B, P1 * K, 1).expand(-1, -1, p2.shape[2])).reshape(B, P1, K, -1)
    d_re = ((p1[:, :, None, :] - nb) ** 2).sum(-1)                          # same values up to summation order
    assert float((d - d_re).abs().max()) <= RTOL * float(d_re.abs().max())
    srt, _ = torch.sort(i, dim=2)
    assert bool((srt[:, :, 1:] != srt[:, :, :-1]).all())                  # K distinct neighbours
    if self_query:
        assert bool((i[:, :, 0] == torch.arange(P1, device=i.device)[None]).all())
        assert float(d[:, :, 0].abs().max()) == 0.0
    # no point outside the list is closer than the K-th neighbour (checked against 64 random candidates)
    g = torch.Generator(device="cuda").manual_seed(0)
    r = torch.randint(0, p2.shape[1], (64,), device="cuda", generator=g)
    d_r = ((p1[:, :, None, :] - p2[:, r][:, None]) ** 2).sum(-1)            # [B,P1,64]
    inlist = (i[:, :, :, None] == r[None, None, None, :]).any(2)
    slack = RTOL * float(d_re.abs().max())
    assert bool(((d_r >= d[:, :, -1:] - slack) | inlist).all())


@pytest.mark.parametrize("D,K", [(32, 20), (64, 12), (32, 9), (64, 8)])
def test_config2_feature_knn_full_size(F, oracle, D, K):
    """configs[1]/[2]: B=8, 2048 x 2048, the generator's (D, K) pairs — the tcgen05 path."""
    rng = np.random.default_rng(100 + D + K)
    x = rng.standard_normal((8, 2048, D)).astype(np.float32)
    xg = cu(x)
    d, i = F.knn(xg, xg, K)
    _check_knn_properties(xg, xg, d, i, K, self_query=True)
    sl = slice(1000, 1096)                                                 # 96 queries per cloud vs all 2048 candidates
    od, oi = oracle.knn(np.ascontiguousarray(x[:, sl]), x, K)
    np.testing.assert_array_equal(i[:, sl].cpu().numpy(), oi)
    np.testing.assert_array_equal(d[:, sl].cpu().numpy(), od)


@pytest.mark.parametrize("N,K", [(8192, 16), (65536, 32)])
def test_config3_3d_knn_and_frnn_full_size(F, oracle, N, K):
    """configs[2] sweep: 3-D kNN / FRNN on N points (uniform-grid search), B=8 for 8192, B=2 for 65536."""
    rng = np.random.default_rng(200 + N)
    B = 8 if N <= 8192 else 2
    p = synth.fluid_cloud(rng, B, N)
    pg = cu(p)
    d, i = F.knn(pg, pg, K)
    _check_knn_properties(pg, pg, d, i, K, self_query=True)
    sl = slice(N // 2, N // 2 + 64)
    od, oi = oracle.knn(np.ascontiguousarray(p[:, sl]), p, K)
    np.testing.assert_array_equal(i[:, sl].cpu().numpy(), oi)
    np.testing.assert_array_equal(d[:, sl].cpu().numpy(), od)
    r = 0.025 * (2 * K * 3 / (4 * np.pi)) ** (1 / 3)                       # ~2K points inside the ball
    fd, fi = F.frnn(pg, pg, K, r)
    ofd, ofi = oracle.frnn(np.ascontiguousarray(p[:, sl]), p, K, r)
    np.testing.assert_array_equal(fi[:, sl].cpu().numpy(), ofi)
    np.testing.assert_array_equal(fd[:, sl].cpu().numpy(), ofd)
    valid = fi >= 0
    assert bool((fd[valid] < r * r).all()) and bool((fd[~valid] == -1).all())
    assert bool(((fi[:, :, 1:] >= 0) <= (fi[:, :, :-1] >= 0)).all())      # padding only at the tail


def test_config3_grouping_fwd_bwd_full_size(F, oracle):
    """configs[2] largest grouping that fits the oracle slice check: C=256, N=M=8192, k=32 (2.2 GB output)."""
    rng = np.random.default_rng(300)
    B, C, N, k = 8, 256, 8192, 32
    f = torch.randn(B, C, N, device="cuda")
    idx = torch.from_numpy(rng.integers(0, N, size=(B, N, k)).astype(np.int32)).cuda()
    out = F.group_fwd(f, idx)
    ref = torch.gather(f[:, :8], 2, idx.long().reshape(B, 1, N * k).expand(-1, 8, -1)).reshape(B, 8, N, k)
    assert torch.equal(out[:, :8], ref)                                    # pure copies: exact
    assert torch.equal(out[:, -1], torch.gather(f[:, -1], 1, idx.long().reshape(B, N * k)).reshape(B, N, k))
    go = torch.randn(B, C, N, k, device="cuda")
    off, items = F.inverse_index(idx, N)
    gf = F.group_bwd(go, off, items, N)
    # scatter-add conserves the per-(b, c) sum of the gradient
    s_in, s_out = go.double().sum((2, 3)), gf.double().sum(2)
    assert float((s_in - s_out).abs().max()) <= 1e-4 * float(go.abs().double().sum((2, 3)).max())
    # two channels of two clouds against the oracle (same summation order -> bit-exact)
    o = oracle.group_bwd(go[:2, :2].cpu().numpy(), idx[:2].cpu().numpy(), N)
    np.testing.assert_array_equal(gf[:2, :2].cpu().numpy(), o)
    del out, go


def test_config2_grouping_backward_generator_shapes(F, oracle):
    """configs[1]: the generator's grouping backward shapes through the staged kernel, real kNN lists."""
    rng = np.random.default_rng(301)
    p = cu(synth.fluid_cloud(rng, 8, 2048))
    for C, k in [(32, 20), (32, 9), (64, 12), (64, 4)]:
        idx = F.knn(p, p, k)[1].to(torch.int32)
        go = torch.randn(8, C, 2048, k, device="cuda")
        off, items = F.inverse_index(idx, 2048)
        gf = F.group_bwd(go, off, items, 2048)
        o = oracle.group_bwd(go[:1, :8].cpu().numpy(), idx[:1].cpu().numpy(), 2048)
        np.testing.assert_array_equal(gf[:1, :8].cpu().numpy(), o)
        ref = torch.zeros(8, C, 2048, device="cuda", dtype=torch.float64)
        ref.scatter_add_(2, idx.long().reshape(8, 1, -1).expand(-1, C, -1), go.double().reshape(8, C, -1))
        assert float((gf.double() - ref).abs().max()) <= RTOL * float(ref.abs().max())


def test_config4_chamfer_full_size(F, oracle):
    """configs[3]: Chamfer forward + backward, src [32,8192,3], tgt [32,32768,3]."""
    rng = np.random.default_rng(400)
    B, P1, P2 = 32, 8192, 32768
    tgt = synth.fluid_cloud(rng, B, P2)
    src = np.ascontiguousarray(tgt[:, ::4] + 0.003 * rng.standard_normal((B, P1, 3)).astype(np.float32))
    s, t = cu(src), cu(tgt)
    r = F.chamfer_fwd(s, t, 3)
    r2 = F.chamfer_fwd(t, s, 3)                                            # swapped clouds: the two sums swap
    np.testing.assert_allclose(r["sum_src"].cpu().numpy(), r2["sum_tgt"].cpu().numpy(), rtol=RTOL)
    np.testing.assert_allclose(r["sum_tgt"].cpu().numpy(), r2["sum_src"].cpu().numpy(), rtol=RTOL)
    assert torch.equal(r["i_src"], r2["i_tgt"]) and torch.equal(r["i_tgt"], r2["i_src"])
    # the nearest target of a source point is no farther than its own parent point tgt[4 i]
    d_nn = ((s - torch.gather(t, 1, r["i_src"].long()[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    d_parent = ((s - t[:, ::4]) ** 2).sum(-1)
    assert bool((d_nn <= d_parent * (1 + 1e-6) + 1e-12).all())
    # two clouds against the oracle: indices exact, sums to 1e-5
    o = oracle.chamfer_fwd(src[:2], tgt[:2], 3)
    np.testing.assert_array_equal(r["i_src"][:2].cpu().numpy(), o["i_src"])
    np.testing.assert_array_equal(r["i_tgt"][:2].cpu().numpy(), o["i_tgt"])
    np.testing.assert_allclose(r["sum_src"][:2].cpu().numpy(), o["sum_src"], rtol=RTOL)
    np.testing.assert_allclose(r["sum_tgt"][:2].cpu().numpy(), o["sum_tgt"], rtol=RTOL)
    g = torch.full((B,), 1.0 / B, device="cuda")
    gs, gt = F.chamfer_bwd(s, t, r["i_src"], r["i_tgt"], g, g, 3)
    # translation invariance: the gradients of both clouds cancel per cloud
    tot = gs.double().sum(1) + gt.double().sum(1)
    assert float(tot.abs().max()) <= 1e-4 * float(gs.abs().double().sum(1).max())
    ogs, ogt = oracle.chamfer_bwd(src[:1], tgt[:1], o["i_src"][:1], o["i_tgt"][:1], np.full((1,), 1.0 / B, np.float32),
                                  np.full((1,), 1.0 / B, np.float32), 3)
    np.testing.assert_allclose(gs[:1].cpu().numpy(), ogs, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(gt[:1].cpu().numpy(), ogt, rtol=RTOL, atol=1e-9)


def test_config2_fps_ball_query_full_size(F, oracle):
    """configs[1]: FPS 8192 -> 1024 and the ball query of the spatial discriminator's first SA level."""
    rng = np.random.default_rng(500)
    xyz = synth.fluid_cloud(rng, 8, 8192)
    xg = cu(xyz)
    idx = F.fps(xg, 1024)
    np.testing.assert_array_equal(idx.cpu().numpy(), oracle.fps(xyz, 1024))
    new_xyz = torch.gather(xg, 1, idx.long()[..., None].expand(-1, -1, 3)).contiguous()
    bq = F.ball_query(0.1, 32, xg, new_xyz)
    o = oracle.ball_query(0.1, 32, xyz, new_xyz.cpu().numpy())
    np.testing.assert_array_equal(bq.cpu().numpy(), o)
