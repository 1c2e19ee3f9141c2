"""``pytorch3d.ops`` call surface used by the reference:
``from pytorch3d.ops import knn_points`` (gcn_lib/pointnet/gcn.py:3, discriminator.py:9,
gcn_lib/interpolation.py:8, gcn_lib/graph_utils.py:3) plus ``knn_gather``.
"""
from collections import namedtuple
from typing import Optional, Union

import torch

from tpugan_b200 import functional as F

_KNN = namedtuple("KNN", "dists idx knn")


class _KnnDists(torch.autograd.Function):
    """dists/idx from the sm_100a kernel; gradient of dists w.r.t. both clouds
    (2 g (p1 - p2[idx]) and its scatter), never taken on the reference's train step
    because no caller consumes `dists` (gcn.py:91,258; discriminator.py:33)."""

    @staticmethod
    def forward(ctx, p1, p2, lengths1, lengths2, K):
        dists, idx = F.knn(p1, p2, K, lengths1, lengths2)
        ctx.save_for_backward(p1, p2, idx)
        ctx.mark_non_differentiable(idx)
        return dists, idx

    @staticmethod
    def backward(ctx, grad_dists, _grad_idx):
        p1, p2, idx = ctx.saved_tensors
        B, P1, K = idx.shape
        D = p1.shape[2]
        nbr = F.gather_rows(p2, idx.reshape(B, P1 * K)).reshape(B, P1, K, D)
        diff = (p1.unsqueeze(2) - nbr) * (2.0 * grad_dists).unsqueeze(-1)  # [B,P1,K,D]
        grad_p1 = diff.sum(2)
        grad_p2 = torch.zeros_like(p2)
        grad_p2.scatter_add_(1, idx.reshape(B, P1 * K, 1).expand(B, P1 * K, D), -diff.reshape(B, P1 * K, D))
        return grad_p1, grad_p2, None, None, None


def knn_points(
    p1: torch.Tensor,
    p2: torch.Tensor,
    lengths1: Union[torch.Tensor, None] = None,
    lengths2: Union[torch.Tensor, None] = None,
    norm: int = 2,
    K: int = 1,
    version: int = -1,
    return_nn: bool = False,
    return_sorted: bool = True,
) -> _KNN:
    """K nearest neighbours of every point of p1 in p2 (squared L2), ascending.

    Same signature and return type as pytorch3d's.  Neighbours are always returned in
    the canonical (distance, index) order, so ``return_sorted=False`` is honoured
    trivially; ``version`` is accepted and ignored.
    """
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if norm != 2:
        raise ValueError("Support for 1 or 2 norm." if norm not in (1, 2) else
                         "tpugan_b200 knn_points implements norm=2 only (the reference never passes norm)")
    p1 = p1.contiguous()
    p2 = p2.contiguous()
    dists, idx = _KnnDists.apply(p1, p2, lengths1, lengths2, int(K))
    nn = None
    if return_nn:
        nn = knn_gather(p2, idx, lengths2)
    return _KNN(dists=dists, idx=idx, knn=nn)


class _KnnGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        N, L, K = idx.shape
        ctx.save_for_backward(idx)
        ctx.M = x.shape[1]
        return F.gather_rows(x, idx.reshape(N, L * K)).reshape(N, L, K, x.shape[2])

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        N, L, K = idx.shape
        U = grad_out.shape[-1]
        gx = torch.zeros((N, ctx.M, U), dtype=grad_out.dtype, device=grad_out.device)
        gx.scatter_add_(1, idx.reshape(N, L * K, 1).expand(N, L * K, U), grad_out.reshape(N, L * K, U))
        return gx, None


def knn_gather(x: torch.Tensor, idx: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,M,U], idx [N,L,K] -> x_out [N,L,K,U] with x_out[n,l,k] = x[n, idx[n,l,k]];
    neighbours beyond lengths[n] are zeroed (pytorch3d semantics)."""
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    out = _KnnGather.apply(x.contiguous(), idx.contiguous())
    if lengths is None:
        if M >= K:
            return out  # nothing to mask, and no host sync
        lengths = torch.full((N,), M, dtype=torch.int64, device=x.device)
    needs_mask = M < K or bool(lengths.min() < K)  # upstream takes the same host read
    if needs_mask:
        mask = lengths[:, None] <= torch.arange(K, device=x.device)[None]  # [N,K]
        mask = mask[:, None].expand(-1, L, -1)[:, :, :, None].expand(-1, -1, -1, U)
        out = out.masked_fill(mask, 0.0)
    return out


__all__ = ["knn_points", "knn_gather"]
