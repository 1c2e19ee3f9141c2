"""GPU replacement for ``sampling.farthest_point_sampling`` (sampling.py:50-106).

Same call and return convention as the reference: NumPy in -> NumPy out
(``indices`` int64 [k], ``distances`` float32 [k, N]).  A CUDA tensor in returns CUDA
tensors and skips the host round trip; a leading batch dimension is accepted in that case.
"""
from __future__ import annotations

import numpy as np
import torch

from . import functional as F


def farthest_point_sampling(pts, k, initial_idx=None, skip_initial=False, indices_dtype=np.int64,
                            distances_dtype=np.float32, return_distances=True, device="cuda"):
    is_numpy = isinstance(pts, np.ndarray)
    if is_numpy:
        assert pts.ndim == 2
        t = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).to(device)[None]
    else:
        t = pts.contiguous().float()
        if t.dim() == 2:
            t = t[None]
    B, N, _ = t.shape
    if initial_idx is None:
        start = torch.from_numpy(np.random.randint(N, size=B).astype(np.int64))  # sampling.py:87-88
    else:
        start = torch.as_tensor(initial_idx, dtype=torch.int64).reshape(-1).expand(B)
    if skip_initial:
        # sampling.py:99-103: replace the start by the point farthest from it
        first = F.fps_start(t, 2, start.to(t.device))
        start = first[:, 1]
    res = F.fps_start(t, int(k), start.to(t.device), return_rows=return_distances)
    idx, rows = res if return_distances else (res, None)
    if is_numpy:
        idx_np = idx[0].cpu().numpy().astype(indices_dtype, copy=False)
        rows_np = rows[0].cpu().numpy().astype(distances_dtype, copy=False) if rows is not None else None
        return idx_np, rows_np
    if pts.dim() == 2:
        return idx[0], (rows[0] if rows is not None else None)
    return idx, rows
