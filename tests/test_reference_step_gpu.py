"""T1: the reference's UNMODIFIED train step (baseline/_ref/train_step_final.py:69-320) and
``SRNet.forward`` (upsampling_network.py:176-185) run on cuda:0 over the drop-in CUDA packages.
Every boundary call the reference makes is recorded (device-side clones of inputs and outputs,
tpugan_b200.recording) and re-checked against the CPU oracle on its own recorded inputs, so the
comparison is exact per call no matter how the dense layers (cuDNN) round.

Covers what only shows up in the real step: positions with requires_grad entering FPS /
ball_query / grouping / gather (train_step_final.py:120,145-150), the (999,999,999) dummy block
and the discriminators' dummy re-draw (upsampling_network.py:143-150, discriminator.py:115-130),
QueryAndGroup inside the set-abstraction layers (discriminator.py:140-148), the in-place kNN fill
of FRNN lists (discriminator.py:39), dtype / contiguity conventions of the reference's casts.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

# boundary calls of one full G+D step (tests/golden/*_step_schedule.json).  The reference's 105 (99)
# grouping_operation calls include the two of every QueryAndGroup (discriminator.py:190), which the drop-in
# pointnet2_utils.QueryAndGroup writes in one assembly pass: verify_calls.grouping_equivalents counts those back.
FLUID_COUNTS = {"knn": 42, "frnn": 11, "chamfer": 1, "fps": 27, "gather": 27, "ball_query": 27,
                "group_bwd": 66, "gather_bwd": 9, "chamfer_bwd": 1, "group_assemble": 27}
FLUID_GROUPINGS = 105
ACTION_COUNTS = {"knn": 36, "frnn": 9, "chamfer": 1, "fps": 27, "gather": 27, "ball_query": 27, "group_assemble": 27}
ACTION_GROUPINGS = 99


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "GPU test selected but no CUDA device"
    return torch


def _snapshot(nets):
    return [[p.detach().clone() for p in m.parameters()] for m in nets]


def _changed(before, nets):
    return [any(not bool((a == p.detach()).all()) for a, p in zip(b, m.parameters())) for b, m in zip(before, nets)]


def test_installed_reference_is_unmodified():
    import install_ref

    assert install_ref.installed(), "baseline/_ref missing: __graft_entry__.build() installs it"
    assert install_ref.verify(), "a file under baseline/_ref differs from the reference it was copied from"


@pytest.mark.parametrize("B,n_lo,ratio", [(2, 512, 4), (3, 1152, 8)])
def test_reference_fluid_step_gpu(torch_cuda, oracle, B, n_lo, ratio):
    import refstep
    import verify_calls
    from tpugan_b200 import launch_count
    from tpugan_b200.recording import log

    ctx = refstep.build("fluid", B=B, n_lo=n_lo, ratio=ratio, backend="cuda")
    assert ctx.mods["ref_dir"].endswith(os.path.join("baseline", "_ref"))
    before = _snapshot(ctx.networks())
    l0 = launch_count()
    log.start(capture=True)
    losses = refstep.step(ctx, n_iter=12)
    calls = log.stop()
    torch_cuda.cuda.synchronize()
    assert launch_count() > l0
    assert all(np.isfinite(v) for v in losses.values()), losses
    assert losses["masking_loss"] < 0.1, "GAN branch (train_step_final.py:117) not taken"
    assert losses["tempo_D_loss"] != 0.0 and losses["spatial_D_loss"] != 0.0, "D updates skipped"
    assert all(_changed(before, ctx.networks())), "an optimiser did not step"
    got = verify_calls.counts(calls)
    extra = {k: v for k, v in got.items() if k not in FLUID_COUNTS}
    assert set(extra) <= {"gather_rows", "group"}, extra  # index_points is plain torch indexing in the reference
    assert {k: got.get(k, 0) for k in FLUID_COUNTS} == FLUID_COUNTS
    assert verify_calls.grouping_equivalents(calls) == FLUID_GROUPINGS
    # the hard-mask path padded with (999,999,999) dummies and FPS met them
    fps_in = [c.inputs["xyz"] for c in calls if c.op == "fps"]
    assert any(bool((x == 999).any()) for x in fps_in), "no dummy block reached FPS"
    seen = verify_calls.check_log(oracle, calls)
    assert sum(n for n, _ in seen.values()) == len(calls)


def test_reference_action_step_gpu(torch_cuda, oracle):
    import refstep
    import verify_calls
    from tpugan_b200.recording import log

    ctx = refstep.build("action", B=2, n_lo=128, ratio=16, backend="cuda")
    before = _snapshot(ctx.networks())
    log.start(capture=True)
    losses = refstep.step(ctx, n_iter=12)
    calls = log.stop()
    torch_cuda.cuda.synchronize()
    assert all(np.isfinite(v) for v in losses.values()), losses
    assert all(_changed(before, ctx.networks()))
    got = verify_calls.counts(calls)
    assert {k: got.get(k, 0) for k in ACTION_COUNTS} == ACTION_COUNTS
    assert verify_calls.grouping_equivalents(calls) == ACTION_GROUPINGS
    assert got.get("group_bwd", 0) > 0 and got.get("chamfer_bwd", 0) == 1
    verify_calls.check_log(oracle, calls)


def test_reference_generator_forward_gpu(torch_cuda, oracle):
    """BASELINE configs[0] shape on the GPU: SRNet(3,128) forward, one 2048-particle frame, batch 1."""
    import refstep
    import verify_calls
    from tpugan_b200.recording import log

    ctx = refstep.build("fluid", B=1, n_lo=2048, ratio=4, backend="cuda")
    log.start(capture=True)
    pos, mask, padded = refstep.generator_forward(ctx)
    calls = log.stop()
    assert pos.shape == (1, 8192, 3) and mask.shape == (1, 2048, 1)
    assert verify_calls.counts(calls) == {"knn": 11, "group": 11}
    verify_calls.check_log(oracle, calls)


def test_reference_step_two_iterations_state_carries(torch_cuda):
    """Odd n_iter = generator-only update (train_step_final.py:166); losses stay finite over steps."""
    import refstep

    ctx = refstep.build("fluid", B=2, n_lo=256, ratio=4, backend="cuda")
    out = [refstep.step(ctx, n_iter=n) for n in (11, 12, 13)]
    assert out[0]["tempo_D_loss"] == 0.0 and out[1]["tempo_D_loss"] != 0.0 and out[2]["tempo_D_loss"] == 0.0
    assert all(np.isfinite(v) for o in out for v in o.values())
