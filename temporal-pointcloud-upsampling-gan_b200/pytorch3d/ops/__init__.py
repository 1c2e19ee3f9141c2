"""``pytorch3d.ops`` call surface used by the reference:
``from pytorch3d.ops import knn_points`` (gcn_lib/pointnet/gcn.py:3, discriminator.py:9,
gcn_lib/interpolation.py:8, gcn_lib/graph_utils.py:3) plus ``knn_gather``.
"""
from collections import namedtuple
from typing import Optional, Union

import torch

from tpugan_b200 import functional as F

_KNN = namedtuple("KNN", "dists idx knn")


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
    dists, idx = F.NeighbourDists.apply(p1, p2, lengths1, lengths2, int(K), None)
    nn = None
    if return_nn:
        nn = knn_gather(p2, idx, lengths2)
    return _KNN(dists=dists, idx=idx, knn=nn)


def knn_gather(x: torch.Tensor, idx: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,M,U], idx [N,L,K] -> x_out [N,L,K,U] with x_out[n,l,k] = x[n, idx[n,l,k]];
    neighbours beyond lengths[n] are zeroed (pytorch3d semantics)."""
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    out = F.GatherRows.apply(x.contiguous(), idx.reshape(N, L * K).contiguous()).reshape(N, L, K, U)
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
