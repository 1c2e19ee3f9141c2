"""``chamferdist.ChamferDistance`` backed by the fused sm_100a nearest-neighbour + reduction
kernels (csrc/chamfer.cu).  Forward value and reduction options follow chamferdist >= 1.0:
per-point nearest squared distance, summed (or averaged) over points, then averaged (or
summed) over the batch; ``bidirectional=True`` adds both directions (loss.py:176-181)."""
import warnings
from typing import Optional

import torch

from tpugan_b200 import functional as F
from tpugan_b200 import _lib


class ChamferDistance(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(
        self,
        source_cloud: torch.Tensor,
        target_cloud: torch.Tensor,
        bidirectional: Optional[bool] = False,
        reverse: Optional[bool] = False,
        batch_reduction: Optional[str] = "mean",
        point_reduction: Optional[str] = "sum",
        reduction: Optional[str] = None,
    ):
        if not isinstance(source_cloud, torch.Tensor):
            raise TypeError("Expected input type torch.Tensor. Got {} instead".format(type(source_cloud)))
        if not isinstance(target_cloud, torch.Tensor):
            raise TypeError("Expected input type torch.Tensor. Got {} instead".format(type(target_cloud)))
        if source_cloud.device != target_cloud.device:
            raise ValueError("Source and target clouds must be on the same device. "
                             f"Got {source_cloud.device} and {target_cloud.device}.")
        if reduction is not None:  # chamferdist 1.0.0 spelling of batch_reduction
            batch_reduction = reduction
        if source_cloud.dim() != 3 or target_cloud.dim() != 3:
            raise ValueError("Expected (B, P, D) point clouds.")
        batchsize_source, lengths_source, dim_source = source_cloud.shape
        batchsize_target, lengths_target, dim_target = target_cloud.shape
        if batchsize_source != batchsize_target:
            raise ValueError("Source and target pointclouds must have the same batchsize.")
        if dim_source != dim_target:
            raise ValueError("Source and target pointclouds must have the same dimensionality.")
        if bidirectional and reverse:
            warnings.warn("Both bidirectional and reverse set to True. bidirectional behavior takes precedence.")
        if point_reduction != "sum" and point_reduction != "mean":
            raise ValueError('Point reduction must either be "sum" or "mean".')
        if batch_reduction != "sum" and batch_reduction != "mean" and batch_reduction is not None:
            raise ValueError('Batch reduction must either be "sum" or "mean".')

        if bidirectional:
            directions = _lib.CHAMFER_BOTH
        elif reverse:
            directions = _lib.CHAMFER_REV
        else:
            directions = _lib.CHAMFER_FWD
        src = source_cloud.contiguous().float()
        tgt = target_cloud.contiguous().float()
        sum_src, sum_tgt = F.ChamferSums.apply(src, tgt, directions)  # [B], [B]

        def _reduce(per_cloud, npoints):
            if point_reduction == "mean":
                per_cloud = per_cloud / npoints
            if batch_reduction == "sum":
                return per_cloud.sum()
            if batch_reduction == "mean":
                return per_cloud.mean()
            return per_cloud

        chamfer_forward = _reduce(sum_src, lengths_source)
        chamfer_backward = _reduce(sum_tgt, lengths_target)
        if bidirectional:
            return chamfer_forward + chamfer_backward
        if reverse:
            return chamfer_backward
        return chamfer_forward
