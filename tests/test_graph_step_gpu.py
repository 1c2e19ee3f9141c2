"""f1: the graph-capturable fluid train step (tpugan_b200.graph_step) equals the reference's unmodified
`tempo_gan_step` given the same seeds — eagerly and as two CUDA graphs — and its device-side dummy re-draw keeps the
reference's semantics."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def _no_dropout(ctx):
    for net in ctx.networks():
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0  # the re-draw consumes the device generator too: keep dropout out of the comparison


def _params(ctx):
    return [torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone() for net in ctx.networks()]


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_graph_safe_step_equals_reference_step():
    import refstep

    torch.backends.cudnn.deterministic = True
    try:
        ctx = refstep.build("fluid", B=2, n_lo=256, ratio=4, backend="cuda", masked_frac=None, capturable=True)
        _no_dropout(ctx)
        snap = refstep.snapshot(ctx)
        p0 = _params(ctx)
        # (1) the reference's own step
        np.random.seed(7)
        torch.manual_seed(7)
        ref_losses = refstep.step(ctx, 12)
        ref_update = [a - b for a, b in zip(_params(ctx), p0)]
        assert ref_losses["masking_loss"] < 0.1 and ref_losses["tempo_D_loss"] != 0.0
        # (2) the host-sync-free restatement, eager
        refstep.restore(ctx, snap)
        gs = refstep.graphed_step(ctx, capture=False)
        np.random.seed(7)
        torch.manual_seed(7)
        got = gs.eager_step(12)
        for k, v in ref_losses.items():
            assert abs(got[k] - v) <= 1e-5 * max(abs(v), 1e-6), (k, got[k], v)
        for name, u, r in zip(("G", "tempoD", "spatialD"), [a - b for a, b in zip(_params(ctx), p0)], ref_update):
            assert _rel(u, r) <= 1e-3, (name, _rel(u, r))
        # (3) as two CUDA graphs; capture (with its warm-up steps) must leave no trace once the state is restored
        refstep.restore(ctx, snap)
        gs = refstep.graphed_step(ctx, capture=True)
        refstep.restore(ctx, snap)
        np.random.seed(7)
        torch.manual_seed(7)
        got2 = gs.step(12)
        for k, v in ref_losses.items():
            assert abs(got2[k] - v) <= 1e-5 * max(abs(v), 1e-6), (k, got2[k], v)
        for name, u, r in zip(("G", "tempoD", "spatialD"), [a - b for a, b in zip(_params(ctx), p0)], ref_update):
            assert _rel(u, r) <= 1e-3, (name, _rel(u, r))
        # odd n_iter: generator update only (train_step_final.py:166)
        got3 = gs.step(13)
        assert got3["tempo_D_loss"] == 0.0 and got3["spatial_D_loss"] == 0.0 and np.isfinite(got3["tempo_G_loss"])
    finally:
        torch.backends.cudnn.deterministic = False


def test_graphed_step_with_dummy_padding_runs_and_trains():
    """per-cloud different keep counts: (999,999,999) padding + device-side re-draw inside the graphs"""
    import refstep

    ctx = refstep.build("fluid", B=3, n_lo=256, ratio=4, backend="cuda", capturable=True)
    gs = refstep.graphed_step(ctx, capture=True)
    p0 = _params(ctx)
    out = [gs.step(n) for n in (12, 13, 14)]
    assert all(np.isfinite(v) for o in out for v in o.values())
    assert out[0]["masking_loss"] < 0.1 and out[0]["tempo_D_loss"] != 0.0 and out[1]["tempo_D_loss"] == 0.0
    assert all(_rel(a, b) > 0 for a, b in zip(_params(ctx), p0))


def test_redraw_dummy_centers_semantics():
    from tpugan_b200.graph_step import redraw_dummy_centers

    g = torch.Generator().manual_seed(3)
    B, N, npnt = 3, 500, 64
    xyz = torch.rand((B, N, 3), generator=g).cuda()
    xyz[0, 100:160] = 999.0
    xyz[2, 5] = 999.0
    ok = torch.tensor([i for i in range(N) if not (100 <= i < 160) and i != 5])
    centers = torch.stack([ok[torch.randperm(len(ok), generator=g)[:npnt]] for _ in range(B)]).to(torch.int32).cuda()
    centers[0, 3], centers[0, 10], centers[0, 40] = 120, 101, 159   # three dummy hits in cloud 0
    centers[2, 63] = 5                                               # one in cloud 2 (cloud 1: none)
    out = redraw_dummy_centers(xyz, centers)
    assert out.dtype == torch.int32 and out.shape == centers.shape
    assert torch.equal(out[1], centers[1])  # untouched cloud
    c0 = centers[0].tolist()
    keep0 = [c for i, c in enumerate(c0) if i not in (3, 10, 40)]
    assert out[0, :61].tolist() == keep0  # survivors first, original order (fps_center[b][~mask[b]], :128)
    fill = out[0, 61:].tolist()
    assert len(set(fill)) == 3 and all(0 <= f < N and f not in (3, 10, 40) for f in fill)  # drawn from [0,N) minus the dropped POSITIONS
    assert out[2, :63].tolist() == centers[2, :63].tolist() and 0 <= int(out[2, 63]) < N and int(out[2, 63]) != 63
