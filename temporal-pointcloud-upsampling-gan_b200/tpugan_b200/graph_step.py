"""Graph-capturable TPU-GAN train step.

The reference's ``tempo_gan_step`` (train_step_final.py:69-230) cannot be captured into a CUDA graph:
it branches on a device scalar (``if ml < 0.1``, :117), draws labels / rotations / permutations on the
host in the middle of the step, reads six losses back with ``.item()`` (:222-229), and the models it
calls synchronise too — the dummy re-draw of the set-abstraction layers (discriminator.py:115-130: a
``torch.any`` + a per-cloud Python loop over ``nonzero`` / ``unique`` with NumPy's RNG), the hard-mask
compaction (upsampling_network.py:143-155: ``torch.any`` + boolean-mask indexing) and the boolean-mask
fill of ``ball_query_wrapper`` (discriminator.py:39).  With ~7 000 kernel launches per step the eager
step is bound by the host, not by the GPU.

This module restates the step with those host round trips removed, **calling the reference's own,
unmodified networks** (``SRNet``, ``FluidSpatialDis``, ``FluidTempoDis`` and their layers):

* every host-side random draw of a step becomes an INPUT (:class:`StepRandoms`, device tensors refreshed
  before each replay).  :func:`draw_randoms` consumes NumPy's / torch's host generators in exactly the
  order the reference does, so with equal seeds both see the same labels, permutations and rotations
  (rotations are always applied; "not drawn" = identity, which is exact in fp32);
* the ``ml < 0.1`` gate multiplies the GAN terms of the generator loss (same value and gradient as the
  reference's branch), and the discriminator phase — which only sees *detached* generator outputs — is a
  second graph that the host launches iff the gate passed: ONE host read per step instead of 8+;
* a handful of functions are rebound while a step is traced (the reference's files stay untouched; also
  ``loss.index_points``, whose batch index is built on the host):
  ``SRNet.expand_pos_with_masking`` -> fixed-shape padding path (``torch.where``; the reference takes this
  path whenever the clouds of a batch keep different numbers of points — with equal counts it compacts
  instead, which changes tensor shapes and cannot be captured),
  ``_PointnetSAModuleBase.forward`` -> device-side dummy re-draw (stable compaction of the surviving FPS
  centres + a random draw without replacement from the allowed indices, all fixed-shape tensor ops with
  the device generator), ``discriminator.ball_query_wrapper`` -> one kNN search, ``IDGCNLayer.forward`` -> one
  neighbour search per layer instead of three + fused gather+max (see reference_patches; identical results).

Equality with the reference step (same seeds, no dummy points in the batch, so that the only random
stream that cannot be shared — NumPy inside the re-draw loop — stays unused) is tested in
tests/test_graph_step_gpu.py: losses and updated weights agree to fp32 rounding.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as TF

from . import gcn_dense

DT = 0.025  # train_step_final.py:7
LOSS_KEYS = ("tempo_G_loss", "tempo_D_loss", "Chamfer_distance_no_norm", "masking_loss", "spatial_G_loss", "spatial_D_loss")


# ------------------------------------------------------------------------------------------ random inputs
@dataclass
class StepRandoms:
    """Everything the reference draws on the host during one step, as device tensors."""
    scalars: torch.Tensor      # [6] valid, invalid, spatial_G target, tempo_G target, (unused), (unused)
    perms: torch.Tensor        # [frames, N_padded] int64: point permutations of the fake frames (centre frame first use: spatial D)
    rot_pred: torch.Tensor     # [frames, 3, 3]   rotation of every fake frame for the tempo-D update (identity if not drawn)
    rot_real: torch.Tensor     # [frames, 3, 3]
    rot_sp_real: torch.Tensor  # [B, 3, 3]  per-cloud rotations of the spatial-D update (identity if not drawn)
    rot_sp_fake: torch.Tensor  # [B, 3, 3]

    def copy_from(self, other: "StepRandoms") -> None:
        for k in ("scalars", "perms", "rot_pred", "rot_real", "rot_sp_real", "rot_sp_fake"):
            getattr(self, k).copy_(getattr(other, k), non_blocking=True)


def _rotation_matrix_np() -> np.ndarray:
    """get_rotation_matrix (train_step_final.py:10-30) on the host, same RNG consumption."""
    a = np.random.uniform(size=3) * 2 * np.pi
    rx = np.array([[1., 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    return (torch.matmul(torch.tensor(rz, dtype=torch.float32),
                         torch.matmul(torch.tensor(ry, dtype=torch.float32), torch.tensor(rx, dtype=torch.float32)))).numpy()


def draw_randoms(frames: int, n_points: int, batch: int, n_iter: int, pinned: bool = True) -> StepRandoms:
    """Host draws of one step in the reference's order (train_step_final.py:85-90,120,122,142,154,170-175,193-207),
    assuming the GAN branch is taken.  Returns HOST tensors (pinned) to be copied into the static device copy."""
    valid = np.random.uniform(0.8, 1.2)
    invalid = np.random.uniform(0.0, 0.2)
    if np.random.uniform(0.0, 1.0) < 0.03:
        valid, invalid = invalid, valid
    perms = torch.empty((frames, n_points), dtype=torch.int64)
    perms[1] = torch.randperm(n_points)                       # :120 spatial D on the permuted centre frame
    sp_target = np.random.uniform(0.8, 1.2)                   # :121
    for f in [0] + list(range(2, frames)):
        perms[f] = torch.randperm(n_points)                   # :142
    tp_target = np.random.uniform(0.8, 1.2)                   # :154
    eye = np.eye(3, dtype=np.float32)
    rot_pred = np.stack([eye] * frames)
    rot_real = np.stack([eye] * frames)
    rot_sp_real = np.stack([eye] * batch)
    rot_sp_fake = np.stack([eye] * batch)
    if n_iter % 2 == 0:
        if np.random.uniform() > 0.7:                         # :170-175 rotate_lst(pred), rotate_lst(real)
            rot_pred = np.stack([_rotation_matrix_np() for _ in range(frames)])
            rot_real = np.stack([_rotation_matrix_np() for _ in range(frames)])
        if np.random.uniform() > 0.7:                         # :193-207
            rot_sp_real = np.stack([_rotation_matrix_np() for _ in range(batch)])
            rot_sp_fake = np.stack([_rotation_matrix_np() for _ in range(batch)])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    r = StepRandoms(scalars=torch.tensor([valid, invalid, sp_target, tp_target, 0.0, 0.0], dtype=torch.float32), perms=perms,
                    rot_pred=t(rot_pred), rot_real=t(rot_real), rot_sp_real=t(rot_sp_real), rot_sp_fake=t(rot_sp_fake))
    if pinned:
        for k in ("scalars", "perms", "rot_pred", "rot_real", "rot_sp_real", "rot_sp_fake"):
            setattr(r, k, getattr(r, k).pin_memory())
    return r


def randoms_like(r: StepRandoms, device) -> StepRandoms:
    return StepRandoms(**{k: torch.empty_like(getattr(r, k), device=device) for k in
                          ("scalars", "perms", "rot_pred", "rot_real", "rot_sp_real", "rot_sp_fake")})


# ------------------------------------------------------------------------------------------ graph-safe model methods
def expand_pos_with_masking_static(self, pos, upsample_edge, binary_mask, hard_masking=False):
    """SRNet.expand_pos_with_masking (upsampling_network.py:131-157), fixed-shape: the padding path
    (`expanded_pos[~hard_mask] = 999`, :149) expressed with torch.where; no host synchronisation."""
    batch_size = pos.shape[0]
    binary_mask = binary_mask.detach()
    binary_mask = binary_mask.view(batch_size, -1, 1) > self.epsilon
    pos_duplicate = torch.cat([pos] * self.upsample_ratio, dim=2)
    upsample_edge = upsample_edge * binary_mask.float()
    expanded_pos = pos_duplicate.view((batch_size, -1, 3)) + upsample_edge.view((batch_size, -1, 3))
    if not hard_masking:
        return expanded_pos, None
    hard_mask = torch.cat([binary_mask] * self.upsample_ratio, dim=2).clone()
    hard_mask[:, :, 0] = True
    hard_mask = hard_mask.view(batch_size, -1, 1)
    unpadded_pos = expanded_pos.clone()
    padded = torch.where(hard_mask, expanded_pos, torch.full_like(expanded_pos, 999.0))
    return unpadded_pos, padded


def redraw_dummy_centers(xyz: torch.Tensor, fps_center: torch.Tensor) -> torch.Tensor:
    """Device-side form of discriminator.py:115-130.  FPS centres that landed on a (999,999,999) dummy are dropped,
    the surviving centres keep their order at the front, and the freed slots are filled with distinct random indices
    drawn (without replacement) from [0, N) minus the *positions* of the dropped entries in the FPS list — the set the
    reference draws from (`combined.unique(...)`, :121-123).  Fixed shapes, device generator, no host round trip."""
    B, N, _ = xyz.shape
    npnt = fps_center.shape[1]
    centers = fps_center.long()
    is_dummy = (torch.gather(xyz[:, :, 0], 1, centers) - 999).abs() < 1e-4            # [B, np]
    n_keep = npnt - is_dummy.sum(1, keepdim=True)                                       # [B, 1]
    order = torch.sort(is_dummy.to(torch.int8), dim=1, stable=True)[1]                  # survivors first, in order
    survivors = torch.gather(centers, 1, order)
    keys = torch.rand((B, N), device=xyz.device)
    banned = torch.zeros((B, N), dtype=torch.bool, device=xyz.device)
    banned[:, :min(npnt, N)] = is_dummy[:, :min(npnt, N)]
    keys = torch.where(banned, torch.full_like(keys, 2.0), keys)
    pool = torch.argsort(keys, dim=1)[:, :npnt]                                         # distinct allowed indices, random order
    slot = torch.arange(npnt, device=xyz.device).unsqueeze(0)
    fill = torch.gather(pool, 1, (slot - n_keep).clamp_min(0))
    return torch.where(slot < n_keep, survivors, fill).to(torch.int32)


def sa_forward_static(self, xyz, features):
    """_PointnetSAModuleBase.forward (discriminator.py:91-153) with the dummy re-draw done on the device."""
    from pointnet2_ops import pointnet2_utils

    new_features_list = []
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    if self.npoint is not None:
        fps_center = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        if self.mask_dummy:
            fps_center = redraw_dummy_centers(xyz, fps_center)
        new_xyz = pointnet2_utils.gather_operation(xyz_flipped, fps_center).transpose(1, 2).contiguous()
    else:
        new_xyz = None
    for i in range(len(self.groupers)):
        new_features = self.groupers[i](xyz, new_xyz, features)
        new_features = self.mlps[i](new_features)
        new_features = TF.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        new_features_list.append(new_features.squeeze(-1))
    return new_xyz, torch.cat(new_features_list, dim=1)


def index_points_static(points, idx):
    """loss.index_points / discriminator.index_points (loss.py:10-27) with the batch index built on the device (the
    reference builds it on the host and copies it over: not capturable)."""
    B = points.shape[0]
    view_shape = [B] + [1] * (idx.dim() - 1)
    batch_indices = torch.arange(B, dtype=torch.long, device=points.device).view(view_shape).expand_as(idx)
    return points[batch_indices, idx, :]


class static_model_methods:
    """Context manager: rebinds the synchronising functions / methods of the reference's modules."""

    def __init__(self, mods: Optional[Dict[str, Any]] = None, fused_idgcn: bool = True, fused_flow: bool = True,
                 restructured_edgeconv: bool = False):
        self.mods = mods or {}
        self.fused_idgcn = fused_idgcn
        self.fused_flow = fused_flow
        self.restructured_edgeconv = restructured_edgeconv
        self._undo: List = []

    def _mod(self, name):
        return self.mods.get(name) or sys.modules.get(name)

    def __enter__(self):
        un, dis, loss = self._mod("upsampling_network"), self._mod("discriminator"), self._mod("loss")
        for obj, name, new in ((un.SRNet, "expand_pos_with_masking", expand_pos_with_masking_static),
                               (dis._PointnetSAModuleBase, "forward", sa_forward_static),
                               (dis, "ball_query_wrapper", gcn_dense.ball_query_wrapper),
                               (dis, "index_points", index_points_static),
                               (loss, "index_points", index_points_static)):
            self._undo.append((obj, name, getattr(obj, name)))
            setattr(obj, name, new)
        if self.fused_idgcn:  # one kNN search per IDGCN layer instead of three + fused gather+max (identical results)
            from .reference_patches import _idgcn_forward_fused

            gcn = sys.modules.get("gcn_lib.pointnet.gcn")
            if gcn is not None:
                self._undo.append((gcn.IDGCNLayer, "forward", gcn.IDGCNLayer.forward))
                gcn.IDGCNLayer.forward = _idgcn_forward_fused
        if self.fused_flow and hasattr(dis, "FlowEmbedding"):  # conv input of the flow embedding in one pass (identical)
            from .reference_patches import _flow_embedding_forward_fused

            self._undo.append((dis.FlowEmbedding, "forward", dis.FlowEmbedding.forward))
            dis.FlowEmbedding.forward = _flow_embedding_forward_fused
        if self.restructured_edgeconv:  # per-node convs + K12 (fp32 reordering: not bit-identical)
            from . import reference_patches as rp

            gcn = sys.modules.get("gcn_lib.pointnet.gcn")
            if gcn is not None:
                if rp._EDGECONV_ORIGINAL[0] is None:
                    rp._EDGECONV_ORIGINAL[0] = gcn.EdgeConv.forward
                self._undo.append((gcn.EdgeConv, "forward", gcn.EdgeConv.forward))
                gcn.EdgeConv.forward = rp._edgeconv_forward_restructured
                self._restore_flag = rp._RESTRUCTURE_EDGECONV
                rp._RESTRUCTURE_EDGECONV = True
        return self

    def __exit__(self, *exc):
        for obj, name, old in reversed(self._undo):
            setattr(obj, name, old)
        self._undo = []
        if self.restructured_edgeconv:
            from . import reference_patches as rp

            rp._RESTRUCTURE_EDGECONV = getattr(self, "_restore_flag", False)
        return False


# ------------------------------------------------------------------------------------------ the step, host-sync free
def generator_phase(mods, sr_net, spatial_dis, tempo_dis, lowres_pos_lst, highres_pos_lst, furthest_distance, opt, n_iter,
                    sr_net_optim, rnd: StepRandoms, side_streams=None):
    """train_step_final.py:92-163 (position-only inputs) without host synchronisation.  Returns the loss tensors and
    the detached fake frames the discriminator phase consumes.  `side_streams` (one per neighbouring frame): the
    generator passes of the neighbouring frames (:128-142) are independent of the centre frame's pass, loss and spatial
    discriminator pass, so they are issued on their own streams (autograd runs their backward there too) and joined
    before the temporal discriminator — same arithmetic, overlapping kernels."""
    tpugan_sr_loss = mods["train_step_final"].tpugan_sr_loss
    lowres = lowres_pos_lst[1]
    others = [0] + list(range(2, len(highres_pos_lst)))
    side = {}
    cur = torch.cuda.current_stream()
    if side_streams:
        for frame, st in zip(others, side_streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                _, _, p = sr_net(lowres_pos_lst[frame], lowres_pos_lst[frame], hard_masking=True)
                side[frame] = (p, p[:, rnd.perms[frame]])
    pred_pos, pred_mask, padded = sr_net(lowres, lowres, hard_masking=True)
    position_loss, cd, ml = tpugan_sr_loss(100., highres_pos_lst[1], pred_pos, lowres, pred_mask,
                                           opt.cutoff / furthest_distance, n_iter)
    gate = (ml.detach() < 0.1).to(torch.float32)  # :117 as a multiplier: same value and gradient as the branch
    fake_label = spatial_dis(padded[:, rnd.perms[1]])
    spatial_loss = (0.5 * (fake_label - rnd.scalars[2]) ** 2).mean()
    pred_lst = [None] * len(highres_pos_lst)
    pred_lst[1] = padded
    last_padded = padded
    for k, frame in enumerate(others):
        if side_streams:
            cur.wait_stream(side_streams[k])
            p, permuted = side[frame]
            p.record_stream(cur)
            permuted.record_stream(cur)
        else:
            _, _, p = sr_net(lowres_pos_lst[frame], lowres_pos_lst[frame], hard_masking=True)
            permuted = p[:, rnd.perms[frame]]
        last_padded = p
        pred_lst[frame] = permuted
    fake_label = tempo_dis(pred_lst, opt.R)
    tempo_loss = (0.5 * (fake_label - rnd.scalars[3]) ** 2).mean()
    sr_loss = gate * (tempo_loss + spatial_loss) + opt.w * position_loss
    sr_net_optim.zero_grad(set_to_none=True)
    sr_loss.backward()
    sr_net_optim.step()
    fakes = [p.detach() for p in pred_lst]
    return dict(tempo_G_loss=(gate * tempo_loss).detach(), spatial_G_loss=(gate * spatial_loss).detach(), cd=cd.detach(),
                ml=ml.detach().reshape(())), fakes, last_padded.detach()


def discriminator_phase(spatial_dis, tempo_dis, fakes, last_padded, highres_pos_lst, opt, tempo_dis_optim, spatial_dis_optim,
                        rnd: StepRandoms):
    """train_step_final.py:166-216 (position-only inputs): both discriminator updates."""
    valid, invalid = rnd.scalars[0], rnd.scalars[1]
    B = highres_pos_lst[1].shape[0]
    pred = [torch.bmm(p, rnd.rot_pred[i].unsqueeze(0).expand(B, 3, 3)) for i, p in enumerate(fakes)]
    real = [torch.bmm(p, rnd.rot_real[i].unsqueeze(0).expand(B, 3, 3)) for i, p in enumerate(highres_pos_lst)]
    fake_label = tempo_dis(pred, opt.R)
    true_label = tempo_dis(real, opt.R)
    tempo_dis_loss = (0.5 * ((true_label - valid) ** 2 + (fake_label - invalid) ** 2)).mean()
    tempo_dis_optim.zero_grad(set_to_none=True)
    tempo_dis_loss.backward()
    tempo_dis_optim.step()
    # the reference rotates `highres_pos_batch`, which rotate_lst(highres_pos_lst) has NOT touched (it rebinds list
    # entries, :42), and the last generated frame (`padded_pred_pos_batch` after the loop of :128-142)
    highres = torch.bmm(highres_pos_lst[1], rnd.rot_sp_real)
    fake = torch.bmm(last_padded, rnd.rot_sp_fake)
    fake_label = spatial_dis(fake)
    true_label = spatial_dis(highres)
    spatial_dis_loss = (0.5 * ((true_label - valid) ** 2 + (fake_label - invalid) ** 2)).mean()
    spatial_dis_optim.zero_grad(set_to_none=True)
    spatial_dis_loss.backward()
    spatial_dis_optim.step()
    return dict(tempo_D_loss=tempo_dis_loss.detach(), spatial_D_loss=spatial_dis_loss.detach())


class GraphedFluidStep:
    """The fluid GAN step as two CUDA graphs (generator phase; discriminator phase) over the reference's networks.

    ``step(n_iter)`` = refresh the random inputs, replay graph 1, read the masking loss (the step's one host
    synchronisation), replay graph 2 iff ``n_iter`` is even and the gate passed, return the six losses of the
    reference (:222-229).  ``lowres`` / ``highres`` are static input buffers: copy new frames into them.
    Optimisers must be capturable (``torch.optim.Adam(..., capturable=True)``)."""

    def __init__(self, mods, sr_net, spatial_dis, tempo_dis, lowres_pos_lst, highres_pos_lst, opt, optims,
                 furthest_distance: float = 1.0, warmup: int = 3, capture: bool = True, overlap_frames: bool = False,
                 fused_idgcn: bool = True, fused_flow: bool = True, restructured_edgeconv: bool = False):
        self.mods, self.nets = mods, (sr_net, spatial_dis, tempo_dis)
        self.lowres = [t.clone() for t in lowres_pos_lst]
        self.highres = [t.clone() for t in highres_pos_lst]
        self.opt, self.fd = opt, furthest_distance
        self.og, self.ot, self.os = optims
        dev = self.lowres[0].device
        B, n_lo, _ = self.lowres[0].shape
        self.frames = len(self.highres)
        self.n_pad = n_lo * sr_net.upsample_ratio
        self._host = draw_randoms(self.frames, self.n_pad, B, 0)
        self.rnd = randoms_like(self._host, dev)
        self.rnd.copy_from(self._host)
        self.captured = False
        self.side_streams = [torch.cuda.Stream() for _ in range(self.frames - 1)] if overlap_frames else None
        self.g_out = self.d_out = None
        self._patch = static_model_methods(mods, fused_idgcn=fused_idgcn, fused_flow=fused_flow,
                                           restructured_edgeconv=restructured_edgeconv)
        self.graph_g = self.graph_d = None
        if capture:
            self._capture(warmup)

    # the two phases on the static buffers
    def _run_g(self):
        out, self.fakes, self.last_padded = generator_phase(self.mods, *self.nets, self.lowres, self.highres, self.fd, self.opt,
                                                            self.n_iter_capture, self.og, self.rnd, self.side_streams)
        return out

    def _run_d(self):
        return discriminator_phase(self.nets[1], self.nets[2], self.fakes, self.last_padded, self.highres, self.opt, self.ot,
                                   self.os, self.rnd)

    def _release_autograd_state(self):
        """torch.nn.utils.spectral_norm leaves `module.weight` = weight_orig / sigma (a non-leaf tensor) on the module
        after every forward; it keeps the autograd graph of that forward alive and with it the AccumulateGrad nodes of
        the parameters, which stay bound to the stream they were first created on (e.g. the default stream of an
        earlier eager step).  A backward captured on another stream would then hop streams and invalidate the capture.
        Detaching the stale attribute frees those nodes; the next forward re-creates them on the current stream."""
        for net in self.nets:
            for m in net.modules():
                w = m.__dict__.get("weight")
                if torch.is_tensor(w) and hasattr(m, "weight_orig") and w.grad_fn is not None:
                    m.weight = w.detach()
            for p in net.parameters():
                p.grad = None

    def _capture(self, warmup):
        from . import functional as Fn

        self.n_iter_capture = 12  # > 10: the masking loss is part of the step (loss.py:171)
        torch.cuda.synchronize()
        self._release_autograd_state()
        side = torch.cuda.Stream()  # warm-up AND capture run on this stream
        side.wait_stream(torch.cuda.current_stream())
        with self._patch, torch.cuda.stream(side):
            for _ in range(warmup):  # allocator warm-up on the capture stream (optimizer state gets created here)
                self._run_g()
                self._run_d()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        Fn.csr_cache.clear()
        self.graph_g, self.graph_d = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with self._patch:
            with torch.cuda.graph(self.graph_g, stream=side):
                self.g_out = self._run_g()
            Fn.csr_cache.clear()
            with torch.cuda.graph(self.graph_d, pool=self.graph_g.pool(), stream=side):
                self.d_out = self._run_d()
        Fn.csr_cache.clear()
        self.captured = True

    def eager_step(self, n_iter: int, randoms: Optional[StepRandoms] = None) -> Dict[str, float]:
        """The same host-sync-free step without graphs (debugging / equality tests)."""
        self.n_iter_capture = n_iter
        self.rnd.copy_from(randoms or draw_randoms(self.frames, self.n_pad, self.lowres[0].shape[0], n_iter))
        with self._patch:
            g = self._run_g()
            ml = float(g["ml"])
            d = self._run_d() if (n_iter % 2 == 0 and ml < 0.1) else None
        return self._losses(g, d)

    def step(self, n_iter: int, randoms: Optional[StepRandoms] = None) -> Dict[str, float]:
        assert self.captured and n_iter > 10, "captured for n_iter > 10 (masking loss active)"
        self.rnd.copy_from(randoms or draw_randoms(self.frames, self.n_pad, self.lowres[0].shape[0], n_iter))
        self.graph_g.replay()
        ml = float(self.g_out["ml"])  # the step's one host synchronisation (the gate of :117 / :166)
        run_d = n_iter % 2 == 0 and ml < 0.1
        if run_d:
            self.graph_d.replay()
        return self._losses(self.g_out, self.d_out if run_d else None)

    @staticmethod
    def _losses(g, d) -> Dict[str, float]:
        vals = torch.stack([g["tempo_G_loss"].reshape(()), (d["tempo_D_loss"] if d else torch.zeros_like(g["ml"])).reshape(()),
                            g["cd"].reshape(()), g["ml"].reshape(()), g["spatial_G_loss"].reshape(()),
                            (d["spatial_D_loss"] if d else torch.zeros_like(g["ml"])).reshape(())]).tolist()
        return dict(zip(LOSS_KEYS, vals))
