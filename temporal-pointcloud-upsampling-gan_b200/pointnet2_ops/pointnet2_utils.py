"""``pointnet2_ops.pointnet2_utils`` call surface, backed by libtpugan_b200.so.

Reference call sites: ``furthest_point_sample`` discriminator.py:114,
``gather_operation`` discriminator.py:132, ``grouping_operation``
gcn_lib/pointnet/gcn.py:207,261 and discriminator.py:270,273, ``QueryAndGroup`` /
``GroupAll`` discriminator.py:189-193.  ``ball_query``, ``three_nn`` and
``three_interpolate`` complete the upstream surface.

As upstream, the lower-case names are ``Function.apply`` aliases, index outputs are
non-differentiable int32, inputs must be contiguous float32 CUDA tensors
(RuntimeError otherwise) and there is no CPU path.
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from tpugan_b200 import functional as F


class FurthestPointSampling(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        r"""xyz (B, N, 3) -> (B, npoint) int32 indices of the sampled points; starts at
        index 0, skips points with |p|^2 <= 1e-3 (upstream quirk), ties -> lowest index."""
        out = F.fps(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        r"""features (B, C, N), idx (B, npoint) int32 -> (B, C, npoint)"""
        return F.GatherOperation.forward(ctx, features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        return F.GatherOperation.backward(ctx, grad_out)


gather_operation = GatherOperation.apply


class ThreeNN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        r"""unknown (B, n, 3), known (B, m, 3) -> dist (B, n, 3) l2 distances, idx (B, n, 3) int32"""
        dist, idx = F.three_nn(unknown, known)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, grad_dist, grad_idx):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        r"""features (B, c, m), idx (B, n, 3) int32, weight (B, n, 3) -> (B, c, n)"""
        return F.ThreeInterpolate.forward(ctx, features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        return F.ThreeInterpolate.backward(ctx, grad_out)


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        r"""features (B, C, N), idx (B, npoint, nsample) int32 -> (B, C, npoint, nsample)"""
        return F.GroupingOperation.forward(ctx, features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        return F.GroupingOperation.backward(ctx, grad_out)


grouping_operation = GroupingOperation.apply


class BallQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        r"""xyz (B, N, 3), new_xyz (B, npoint, 3) -> (B, npoint, nsample) int32"""
        out = F.ball_query(radius, nsample, xyz, new_xyz)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    r"""Groups with a ball query of radius

    Parameters
    ----------
    radius : float32
    nsample : int32
        Maximum number of features to gather in the ball
    """

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None):
        r"""xyz (B, N, 3), new_xyz (B, npoint, 3), features (B, C, N) -> (B, 3 + C, npoint, nsample)"""
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        centre = new_xyz.transpose(1, 2).contiguous()
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        # one pass writes cat([xyz[idx] - new_xyz, features[idx]]) (K11, tpg_group_assemble_f32): no per-tensor
        # [B,C,npoint,nsample] intermediates, no separate "- centre" pass, no torch.cat copy; same values
        modes, tensors = [], []
        if self.use_xyz or features is None:
            modes.append("gather")
            tensors += [xyz_trans, centre]
        if features is not None:
            modes.append("gather")
            tensors += [features.contiguous(), None]
        return F.GroupAssemble.apply(idx, tuple(modes), *tensors)


class GroupAll(nn.Module):
    r"""Groups all features"""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None):
        r"""xyz (B, N, 3), new_xyz ignored, features (B, C, N) -> (B, C + 3, 1, N)"""
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)  # (B, 3 + C, 1, N)
            else:
                new_features = grouped_features
        else:
            new_features = grouped_xyz
        return new_features
