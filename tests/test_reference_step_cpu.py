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


@needs_ref
def test_patch_logic_on_cpu_edgeconv_restructure_and_flow_assembly(oracle, monkeypatch):
    """Host logic of tpugan_b200.reference_patches (no CUDA): with the two fused kernels replaced by plain-torch
    stand-ins, the restructured EdgeConv and the one-pass FlowEmbedding input equal the reference's own layers (fp32,
    1e-5: the oracle shims group in fp32) — the algebra W(f_j - f_i) = W f_j - W f_i, the bias / centre handling, the spectral-norm call count and
    the eligibility test (a norm layer inside the affine branches keeps the reference path)."""
    import torch
    import torch.nn.functional as Fn

    import refstep

    mods = refstep.import_reference("oracle")
    dis = mods["discriminator"]
    gcn = sys.modules["gcn_lib.pointnet.gcn"]
    sys.path.insert(0, os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200"))
    from tpugan_b200 import functional as F
    from tpugan_b200 import reference_patches as rp

    def gather(f, idx):
        return torch.stack([f[b][:, idx[b].long()] for b in range(f.shape[0])])

    class EdgeAffineTorch:
        @staticmethod
        def apply(p, q, center, idx, slope):
            return gather(p, idx) + Fn.leaky_relu(gather(q, idx) - center.unsqueeze(-1), slope)

    class GroupAssembleTorch:
        @staticmethod
        def apply(idx, modes, *t):
            outs = []
            for n, m in enumerate(modes):
                src, cen = t[2 * n], t[2 * n + 1]
                if m == "gather":
                    g = gather(src, idx)
                    outs.append(g if cen is None else g - cen.unsqueeze(-1))
                else:
                    outs.append(src.unsqueeze(-1).expand(-1, -1, -1, idx.shape[2]))
            return torch.cat(outs, dim=1)

    monkeypatch.setattr(F, "EdgeAffine", EdgeAffineTorch)
    monkeypatch.setattr(F, "GroupAssemble", GroupAssembleTorch)
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    B, N = 2, 120
    x = torch.from_numpy(rng.standard_normal((B, 8, N))).float()
    edge = gcn.EdgeConv(8, 12, k=8, dilation=2, aggregate='max', mlp_layer=True, bn=False, insn=False)
    edge1 = gcn.EdgeConv(8, 12, k=6, dilation=1, aggregate='max', mlp_layer=False, bn=False, insn=False)
    edge_bn = gcn.EdgeConv(8, 12, k=4, bn=True)
    assert rp.edgeconv_restructurable(edge) and rp.edgeconv_restructurable(edge1) and not rp.edgeconv_restructurable(edge_bn)
    flow = dis.FlowEmbedding(5, [6, 7], sn=False)
    pos1 = torch.from_numpy(rng.uniform(-0.1, 0.1, (B, 3, N))).float()
    pos2 = torch.from_numpy(rng.uniform(-0.1, 0.1, (B, 3, N))).float()
    f1, f2 = (torch.from_numpy(rng.standard_normal((B, 5, N))).float() for _ in range(2))
    ref = [edge(x), edge1(x), edge_bn(x), flow(pos1, pos2, f1, f2, 0.05)[1]]
    h = rp.patch_reference(mods, ball_query=False, idgcn=False, interpolation_kernel=False, flow_embedding=True, edgeconv=True)
    try:
        got = [edge(x), edge1(x), edge_bn(x), flow(pos1, pos2, f1, f2, 0.05)[1]]
    finally:
        h.unpatch()
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert float((a - b).detach().abs().max()) <= 2e-5 * max(float(b.detach().abs().max()), 1.0)
    assert gcn.EdgeConv.forward.__name__ == "forward" and not rp._RESTRUCTURE_EDGECONV
