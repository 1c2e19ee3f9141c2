"""Drop-in for ``frnn`` (lxxue/FRNN) as imported by the reference: ``import frnn`` then
``frnn.frnn_grid_points(...)`` — discriminator.py:27, loss.py:105,142,229,256,261,
gcn_lib/interpolation.py:20,33, gcn_lib/graph_utils.py:46, gcn_lib/pointnet/gcn.py:30.
"""
from typing import Optional, Union

import torch

from tpugan_b200 import functional as F


def frnn_gather(x: torch.Tensor, idxs: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,P2,D], idxs [N,P1,K] (-1 padded) -> [N,P1,K,D]; padded slots are zero."""
    N, P1, K = idxs.shape
    valid = idxs >= 0
    safe = torch.where(valid, idxs, torch.zeros_like(idxs)).contiguous()
    # differentiable w.r.t. x (upstream's frnn_gather is); padded slots give zero rows and take no gradient
    out = F.GatherRows.apply(x.contiguous(), torch.where(valid, safe, torch.full_like(safe, -1)).reshape(N, P1 * K))
    out = out.reshape(N, P1, K, x.shape[2])
    return out * valid.unsqueeze(-1).to(out.dtype)


def frnn_grid_points(
    points1: torch.Tensor,
    points2: torch.Tensor,
    lengths1: Union[torch.Tensor, None] = None,
    lengths2: Union[torch.Tensor, None] = None,
    K: int = -1,
    r: Union[float, torch.Tensor] = -1,
    grid=None,
    return_nn: bool = True,
    return_sorted: bool = True,
    radius_cell_ratio: float = 2.0,
):
    """K nearest neighbours of points1 in points2 within radius r (strict d^2 < r^2).

    Returns ``(dists, idxs, nn, grid)`` like upstream: squared distances and int64
    indices padded with -1, always sorted by (distance, index).  ``grid`` is accepted
    for signature compatibility; the search structure is rebuilt per call on the
    device without host synchronisation, so ``None`` is returned for it.
    """
    if not (isinstance(points1, torch.Tensor) and isinstance(points2, torch.Tensor)):
        raise TypeError("points1 and points2 must be torch.Tensor")
    if not (points1.is_cuda and points2.is_cuda):
        raise TypeError("for now only cuda version is supported")
    if points1.shape[0] != points2.shape[0]:
        raise ValueError("points1 and points2 must have the same batch dimension")
    if points1.shape[2] != points2.shape[2]:
        raise ValueError("dimension mismatch")
    if K <= 0:
        raise ValueError("K must be a positive integer")
    if not isinstance(r, torch.Tensor) and r <= 0:
        raise ValueError("r must be positive")
    # dists are differentiable w.r.t. both clouds like upstream's (padded slots carry no gradient)
    dists, idxs = F.NeighbourDists.apply(points1.contiguous().float(), points2.contiguous().float(), lengths1, lengths2,
                                         int(K), r)
    nn = frnn_gather(points2, idxs, lengths2) if return_nn else None
    return dists, idxs, nn, None


__all__ = ["frnn_grid_points", "frnn_gather"]
