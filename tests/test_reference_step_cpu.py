"""CPU side of T1: the installed reference copy is byte-identical to its manifest, and the
reference's unmodified train steps run over the oracle shims through the same harness the GPU
test uses (tools/refstep.py), making the boundary calls SURVEY.md §3.1 lists."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import install_ref  # noqa: E402

needs_ref = pytest.mark.skipif(not (install_ref.installed() or os.path.isdir("/root/reference")),
                               reason="reference not installed (run __graft_entry__.build() in the build container)")


@needs_ref
def test_installed_copy_matches_manifest_and_source():
    if not install_ref.installed():
        install_ref.install()
    assert install_ref.verify()
    if os.path.isdir(install_ref.SRC):  # build container: compare with the read-only source tree too
        import json

        with open(install_ref.MANIFEST) as f:
            manifest = json.load(f)
        assert len(manifest) >= 20
        for rel, h in manifest.items():
            assert install_ref._sha(os.path.join(install_ref.SRC, rel)) == h, rel


@needs_ref
@pytest.mark.parametrize("domain,n_lo,ratio,want", [
    ("fluid", 128, 4, {"knn": 42, "group": 105, "frnn": 11, "chamfer": 1, "fps": 27, "gather": 27, "ball_query": 27,
                       "group_bwd": 66, "gather_bwd": 9, "chamfer_bwd": 1}),
    ("action", 64, 16, {"knn": 36, "group": 99, "frnn": 9, "chamfer": 1, "fps": 27, "gather": 27, "ball_query": 27}),
])
def test_reference_step_over_oracle_shims(oracle, domain, n_lo, ratio, want):
    import refstep
    import oracle.shims as sh

    ctx = refstep.build(domain, B=2, n_lo=n_lo, ratio=ratio, backend="oracle")
    sh.recorder.start(shapes_only=True)
    losses = refstep.step(ctx, n_iter=12)
    calls = sh.recorder.stop()
    assert all(np.isfinite(v) for v in losses.values()), losses
    if domain == "fluid":
        assert losses["masking_loss"] < 0.1 and losses["tempo_D_loss"] != 0.0
    got = {}
    for op, _, _ in calls:
        got[op] = got.get(op, 0) + 1
    assert {k: got.get(k, 0) for k in want} == want
