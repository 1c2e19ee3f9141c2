"""Harness that runs the reference's UNMODIFIED train step (``baseline/_ref/train_step_final.py``)
on top of either boundary implementation:

* ``backend="cuda"``   — the drop-in packages of this repo (``pytorch3d.ops``, ``frnn``,
  ``pointnet2_ops``, ``chamferdist`` backed by libtpugan_b200.so), everything on ``cuda:0``;
* ``backend="oracle"`` — the CPU oracle shims (``oracle/shims``), for CPU tests and as the
  recorded-schedule source.

Test / bench infrastructure: it lives outside the product package and is the only code that
imports the reference.  The reference files are never edited: the harness only (a) puts the
boundary packages and import-only stubs on ``sys.path``, (b) builds the reference's own models
and optimisers, (c) feeds synthetic frames, and (d) optionally registers a
forward hook on ``SRNet.filter_block`` that gives the mask head the *values* of a trained one
(1 for kept points, 0 for a few percent per cloud; gradients pass straight through), because
the GAN branch of ``tempo_gan_step`` is gated on ``masking_loss < 0.1``
(train_step_final.py:117) which a random-init generator never reaches.  With per-cloud
different keep counts the hard-mask path pads with (999,999,999) dummies
(upsampling_network.py:143-150) and the discriminators' dummy re-draw runs
(discriminator.py:115-130) — the paths a trained model exercises.
"""
from __future__ import annotations

import os
import sys
import warnings
from argparse import Namespace
from typing import Any, Dict, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SITE = os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200")
REF_INSTALLED = os.path.join(ROOT, "baseline", "_ref")
REF_SOURCE = "/root/reference"
BASE_RADIUS = 0.025  # train_utils.py:10
_REF_MODULES = ("train_step_final", "loss", "train_utils", "gcn_lib", "gcn_lib.nn", "gcn_lib.graph_utils", "gcn_lib.gcn",
                "gcn_lib.interpolation", "gcn_lib.pointnet", "gcn_lib.pointnet.gcn", "upsampling_network",
                "discriminator", "sampling", "utils")
_BOUNDARY_MODULES = ("pytorch3d", "pytorch3d.ops", "frnn", "pointnet2_ops", "pointnet2_ops.pointnet2_utils",
                     "chamferdist", "chamferdist.chamfer", "dgl", "dgl.utils", "dgl.function", "dgl.nn", "dgl.geometry",
                     "emd", "open3d", "tensorboardX")


def reference_dir() -> str:
    """The installed copy (ships to the GPU box); falls back to the read-only source tree in
    the build container when ``tools/install_ref.py`` has not been run."""
    if os.path.exists(os.path.join(REF_INSTALLED, "train_step_final.py")):
        return REF_INSTALLED
    if os.path.exists(os.path.join(REF_SOURCE, "train_step_final.py")):
        return REF_SOURCE
    raise FileNotFoundError("reference not installed: run `python tools/install_ref.py` in the build container "
                            "(baseline/_ref is git-ignored and ships with gpurun)")


def import_reference(backend: str = "cuda") -> Dict[str, Any]:
    """Import the reference's modules over the chosen boundary implementation.  Re-importable:
    switching backend purges the cached reference and boundary modules first."""
    assert backend in ("cuda", "oracle")
    for name in _REF_MODULES + _BOUNDARY_MODULES:
        sys.modules.pop(name, None)
    ref = reference_dir()
    stubs = os.path.join(SITE, "import_stubs")
    shim_dir = os.path.join(ROOT, "oracle", "shims")
    for p in (ref, stubs, SITE, shim_dir, ROOT):
        while p in sys.path:
            sys.path.remove(p)
    if backend == "cuda":
        order = [SITE, stubs, ref]
    else:
        sys.path.insert(0, ROOT)
        import oracle.shims  # noqa: F401  (test infrastructure)

        order = [shim_dir, stubs, ref]
        import torch

        if not torch.cuda.is_available():
            # the reference hard-codes .cuda() (train_step_final.py:30,156, loss.py:174)
            torch.Tensor.cuda = lambda self, *a, **k: self
    for p in reversed(order):
        sys.path.insert(0, p)
    sys.path.insert(0, ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SyntaxWarning: `is 'concat'` (discriminator.py:242)
        import discriminator
        import loss
        import train_step_final
        import upsampling_network
    return dict(train_step_final=train_step_final, loss=loss, discriminator=discriminator,
                upsampling_network=upsampling_network, backend=backend, ref_dir=ref)


# --------------------------------------------------------------------------------- synthetic data
def fluid_frames(rng, B, n_hi, frames=3):
    """Fluid-like window: uniform cloud at SPH spacing 0.025, centroid-centred (train_utils.py:214-221);
    neighbouring frames = positions advected by vel*DT with vel ~ 0.5 N(0,1) (train_step_final.py:7,33-35)."""
    L = BASE_RADIUS * (n_hi ** (1.0 / 3.0))
    p = rng.uniform(0.0, L, size=(B, n_hi, 3)).astype(np.float32)
    p -= p.mean(axis=1, keepdims=True).astype(np.float32)
    vel = (0.5 * rng.standard_normal((B, n_hi, 3))).astype(np.float32)
    out = []
    for f in range(frames):
        out.append(np.ascontiguousarray(p + np.float32((f - frames // 2) * 0.025) * vel, dtype=np.float32))
    return out, vel


def action_frames(rng, B, n_hi, frames=3):
    """MSR-Action-like box [-.5,.5]x[-1,1]x[-.25,.25] (msr_dataset.py:81-84), small frame-to-frame motion."""
    lo = np.array([-0.5, -1.0, -0.25], np.float32)
    p = (rng.uniform(size=(B, n_hi, 3)).astype(np.float32) * (-2 * lo) + lo).astype(np.float32)
    out = []
    for f in range(frames):
        out.append(np.ascontiguousarray(p + np.float32(0.01 * f) * rng.standard_normal((B, n_hi, 3)).astype(np.float32)))
    return out, None


class StepContext:
    """Models, optimisers, frames and options of one reference train-step configuration."""

    def __init__(self, domain, mods, sr_net, spatial_dis, tempo_dis, optims, lo, hi, vel_lo, vel_hi, opt, device):
        self.domain, self.mods = domain, mods
        self.sr_net, self.spatial_dis, self.tempo_dis = sr_net, spatial_dis, tempo_dis
        self.optims = optims  # (G, tempo-D, spatial-D)
        self.lo, self.hi, self.vel_lo, self.vel_hi = lo, hi, vel_lo, vel_hi
        self.opt, self.device = opt, device
        self.hook = None
        self.keep = None

    def networks(self):
        return self.sr_net, self.tempo_dis, self.spatial_dis


def build(domain: str = "fluid", B: int = 2, n_lo: int = 512, ratio: int = 4, backend: str = "cuda", device=None,
          seed: int = 1, trained_mask: bool = True, masked_frac=(0.01, 0.06), use_vel: bool = False,
          lr: float = 1e-4, mods: Optional[Dict[str, Any]] = None, capturable: bool = False) -> StepContext:
    """Reference models + synthetic frames.  Shapes of BASELINE configs[1]: B=8, n_lo=2048, ratio=4."""
    import torch

    mods = mods or import_reference(backend)
    if device is None:
        device = torch.device("cuda:0" if backend == "cuda" else "cpu")
    device = torch.device(device)
    torch.manual_seed(seed)  # the reference seeds np / torch / cuda with 1 (train_tempo.py:24-26)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    n_hi = n_lo * ratio
    un, dis = mods["upsampling_network"], mods["discriminator"]
    if domain == "fluid":
        hi_np, vel_np = fluid_frames(rng, B, n_hi)
        in_feats = 6 if use_vel == 6 else 3
        g = un.SRNet(in_feats, 128, upsample_ratio=ratio)
        sd, td = dis.FluidSpatialDis(), dis.FluidTempoDis(3)
        opt = Namespace(use_vel=bool(use_vel), in_node_feats=in_feats, R=0.10, cutoff=0.025, w=0.5)
    elif domain == "action":
        hi_np, vel_np = action_frames(rng, B, n_hi)
        g = un.NoMaskSRNet(3, 128, ratio)
        sd, td = dis.ActionSpatialDis(), dis.ActionTempoDis(3)
        opt = Namespace(R=2.0, w=2.0)
    else:
        raise ValueError(domain)
    g, sd, td = g.to(device), sd.to(device), td.to(device)
    hi = [torch.from_numpy(h).to(device) for h in hi_np]
    # low-res = every ratio-th particle + N(0, 0.003^2) jitter (tempo_dataset.py:27,92; the dataset's FPS
    # down-sampling is exercised separately, tests/test_parity_gpu.py)
    jitter = 0.003 if domain == "fluid" else 0.0
    lo = [(h[:, ::ratio] + jitter * torch.from_numpy(rng.standard_normal((B, n_lo, 3)).astype(np.float32)).to(device)
           ).contiguous() for h in hi]
    vel_hi = vel_lo = None
    if domain == "fluid" and vel_np is not None:
        v = torch.from_numpy(vel_np).to(device)
        vel_hi = [v.clone() for _ in hi]
        vel_lo = [v[:, ::ratio].contiguous() for _ in hi]
    # capturable: Adam keeps its step count on the device (required inside CUDA-graph capture; same update rule)
    optims = tuple(torch.optim.Adam(m.parameters(), lr=lr, capturable=capturable) for m in (g, td, sd))
    ctx = StepContext(domain, mods, g, sd, td, optims, lo, hi, vel_lo, vel_hi, opt, device)
    if domain == "fluid" and trained_mask:
        # values of a trained mask head: 1 = keep, 0 = drop; 1-6 % dropped, a different count per cloud
        keep = np.ones((B, n_lo, 1), np.float32)
        if masked_frac is not None:  # None: every point kept (no dummy padding anywhere in the step)
            fr = np.linspace(masked_frac[0], masked_frac[1], B) if B > 1 else np.array([masked_frac[1]])
            for b in range(B):
                keep[b, rng.choice(n_lo, size=max(1, int(round(fr[b] * n_lo))), replace=False)] = 0.0
        ctx.keep = torch.from_numpy(keep).to(device)

        def trained_mask_hook(_module, _inputs, out):
            k = ctx.keep
            if out.shape != k.shape:
                return out
            return k + (out - out.detach())  # straight-through: trained values, untouched gradient path

        ctx.hook = g.filter_block.register_forward_hook(trained_mask_hook)
    return ctx


def step(ctx: StepContext, n_iter: int = 12, freeze_D: bool = False) -> Dict[str, float]:
    """One call of the reference's train step (G update + both D updates when n_iter is even)."""
    tsf = ctx.mods["train_step_final"]
    og, ot, os_ = ctx.optims
    hi = list(ctx.hi)  # the step rebinds list entries when it draws the rotation augmentation (:172-175)
    lo = list(ctx.lo)
    if ctx.domain == "fluid":
        return tsf.tempo_gan_step(ctx.sr_net, ctx.spatial_dis, ctx.tempo_dis, lo, ctx.vel_lo, hi, ctx.vel_hi, 1.0,
                                  ctx.opt, n_iter, og, ot, os_, freeze_D=freeze_D)
    return tsf.tempo_gan_step_no_mask(ctx.sr_net, ctx.spatial_dis, ctx.tempo_dis, lo, hi, ctx.opt, n_iter, og, ot,
                                      os_, freeze_D=freeze_D)


def graphed_step(ctx: StepContext, capture: bool = True, warmup: int = 3, overlap_frames: bool = False,
                 fused_idgcn: bool = True, restructured_edgeconv: bool = False):
    """The graph-capturable form of the fluid step (tpugan_b200.graph_step) over this context's networks, frames and
    (capturable) optimisers."""
    from tpugan_b200.graph_step import GraphedFluidStep

    assert ctx.domain == "fluid"
    og, ot, os_ = ctx.optims
    return GraphedFluidStep(ctx.mods, ctx.sr_net, ctx.spatial_dis, ctx.tempo_dis, ctx.lo, ctx.hi, ctx.opt, (og, ot, os_),
                            furthest_distance=1.0, warmup=warmup, capture=capture, overlap_frames=overlap_frames,
                            fused_idgcn=fused_idgcn, restructured_edgeconv=restructured_edgeconv)


def snapshot(ctx: StepContext):
    """Deep copy of every parameter / buffer / optimiser-state tensor (restore() copies back IN PLACE, so CUDA graphs
    captured over these tensors stay valid)."""
    import torch

    nets = [{k: v.detach().clone() for k, v in net.state_dict().items()} for net in ctx.networks()]
    opts = [[{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for st in o.state.values()]
            for o in ctx.optims]
    return nets, opts


def restore(ctx: StepContext, snap, zero_new_optimizer_state: bool = True):
    import torch

    nets, opts = snap
    with torch.no_grad():
        for net, st in zip(ctx.networks(), nets):
            cur = net.state_dict()
            for k, v in st.items():
                cur[k].copy_(v)
        for o, saved in zip(ctx.optims, opts):
            for i, st in enumerate(o.state.values()):
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if i < len(saved) and k in saved[i]:
                            v.copy_(saved[i][k])
                        elif zero_new_optimizer_state:
                            v.zero_()


def generator_forward(ctx: StepContext):
    """BASELINE configs[0]: SRNet.forward on the centre frame (upsampling_network.py:176-185)."""
    import torch

    with torch.no_grad():
        if ctx.domain == "fluid":
            return ctx.sr_net(ctx.lo[1], ctx.lo[1], hard_masking=True)
        return ctx.sr_net(ctx.lo[1], ctx.lo[1])


def param_counts(ctx: StepContext):
    return tuple(sum(p.numel() for p in m.parameters() if p.requires_grad) for m in ctx.networks())


# --------------------------------------------------------------------------------- data parallel
class DataParallel:
    """Batch-sharded replicas of the reference's train step WITHOUT touching its code:

    * gradients: an optimiser pre-step hook packs the gradients of that optimiser's network into ONE flat
      fp32 bucket, all-reduces it (NCCL over NVLink; gloo in the CPU tests) and unpacks the mean — three
      bucket all-reduces per GAN step (G, tempo-D, spatial-D; train_step_final.py:161-163,188-190,214-216);
    * branch flag: `tpugan_sr_loss` is wrapped so the masking loss that gates the GAN branch
      (train_step_final.py:117) is the mean over ranks — every rank takes the same branch;
    * BatchNorm statistics (discriminator.py:74-75,357-361): `sync_bn=True` converts the discriminators'
      BatchNorm layers to torch's SyncBatchNorm (exact single-process statistics, CUDA only); otherwise
      statistics stay per-rank like torch DDP's default (documented divergence).
    """

    def __init__(self, ctx: StepContext, sync_bn: bool = False):
        import torch
        import torch.distributed as dist

        self.ctx, self.dist, self.torch = ctx, dist, torch
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.buckets = []
        self.bytes_per_step = 0
        if self.world == 1:
            return
        if sync_bn:
            ctx.spatial_dis = torch.nn.SyncBatchNorm.convert_sync_batchnorm(ctx.spatial_dis)
            ctx.tempo_dis = torch.nn.SyncBatchNorm.convert_sync_batchnorm(ctx.tempo_dis)
        for net in ctx.networks():  # identical replicas: rank 0's initial weights everywhere
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t.data, src=0)
        for net, optim in zip(ctx.networks(), ctx.optims):
            params = [p for p in net.parameters() if p.requires_grad]
            flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=ctx.device)
            self.buckets.append(flat)
            optim.register_step_pre_hook(self._make_hook(params, flat))
        tsf = ctx.mods["train_step_final"]
        orig = tsf.tpugan_sr_loss
        world = self.world

        def agreed_sr_loss(*a, **k):
            loss, cd, ml = orig(*a, **k)
            ml_all = ml.detach().clone().float().reshape(-1)
            dist.all_reduce(ml_all, op=dist.ReduceOp.SUM)
            # value agreed over ranks (same branch everywhere), gradient of the local term untouched
            return loss, cd, ml + (ml_all.reshape(ml.shape) / world - ml.detach())

        tsf.tpugan_sr_loss = agreed_sr_loss

    def _make_hook(self, params, flat):
        dist, world, torch = self.dist, self.world, self.torch

        def hook(_optim, _args, _kwargs):
            o = 0
            for p in params:
                n = p.numel()
                if p.grad is not None:
                    flat[o:o + n].copy_(p.grad.reshape(-1))
                else:
                    flat[o:o + n].zero_()
                o += n
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(world)
            o = 0
            for p in params:
                n = p.numel()
                if p.grad is not None:
                    p.grad.copy_(flat[o:o + n].view_as(p.grad))
                o += n
            self.bytes_per_step += flat.numel() * 4

        return hook
