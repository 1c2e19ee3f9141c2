"""GPU form of the reference's per-sample data preparation.

Reference (CPU, in DataLoader workers): ``FluidTempoDataset.__getitem__`` (train_fluid/tempo_dataset.py:58-105)
-> ``normalize_point_cloud`` (train_utils.py:214-221) -> ``sample_patch_with_fps`` (train_utils.py:98-139: KD-tree
query of the ``patch_num`` points nearest to a random seed particle, numba FPS of the patch to 1/8 —
``sampling.farthest_point_sampling``, sampling.py:50-106) -> index the neighbouring frames with the same patch /
FPS indices -> Gaussian jitter on the low-resolution copies.  The FPS alone costs 0.09 s per 9216 -> 1152 sample on
a host core, which starves 8 GPUs.

Here the whole window is prepared on the device: the frames' ``pos`` / ``vel`` arrays (npz format of
fluid_data_generation/process_training_data.py:75-79) are uploaded once, the patch is the ``patch_num`` smallest
squared distances to the seed in (d2, index) order (what ``KDTree.query`` returns, ties aside), the FPS is the
sampling.py mode of the sm_100a kernel (``tpg_fps_start_f32``: explicit start, first-max arg-max, int64), and the
jitter comes from the device generator.  The random choices (seed particle, FPS start) are arguments, so a caller can
reproduce the reference's NumPy stream.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import functional as F

BASE_RADIUS = 0.025  # train_utils.py:10


def patch_num_for(total_num: int, sample_num: Optional[int]) -> int:
    """train_utils.py:106-117"""
    if sample_num is None:
        return 9216 if total_num > 10000 else int(total_num // 1024) * 1024
    return sample_num if total_num > sample_num else 4096


def sample_patch_with_fps(pos: torch.Tensor, sample_num: Optional[int] = None, seed_idx: Optional[int] = None,
                          fps_start: Optional[int] = None, ds_ratio: float = 0.125):
    """pos [N,3] (CUDA) -> (patch_idx int64 [patch_num], ascending distance from the seed particle; fps_idx int64
    [int(ds_ratio * patch_num)] into the patch).  train_utils.sample_patch_with_fps without the KD-tree."""
    assert pos.is_cuda and pos.dim() == 2
    total = pos.shape[0]
    patch_num = min(patch_num_for(total, sample_num), total)
    if seed_idx is None:
        seed_idx = int(np.random.choice(total))                 # train_utils.py:121
    d = pos.double() - pos[seed_idx].double()   # float64 like scipy's KD-tree, so near-ties order the same way
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    patch_idx = torch.sort(d2, stable=True)[1][:patch_num]      # the patch_num nearest, (d2, index) order
    patch_pos = pos[patch_idx].contiguous()
    if fps_start is None:
        fps_start = int(np.random.randint(patch_num))            # sampling.py:87-88
    fps_idx = F.fps_start(patch_pos[None], int(ds_ratio * patch_num), torch.tensor([fps_start]))[0]
    return patch_idx, fps_idx


def fluid_window(frames: Sequence[Dict[str, torch.Tensor]], sample_num: int = 4096, jitter: float = 0.003,
                 seed_idx: Optional[int] = None, fps_start: Optional[int] = None, generator=None) -> Dict[str, torch.Tensor]:
    """FluidTempoDataset.__getitem__ (tempo_dataset.py:58-105) for one window of (left, centre, right) frames, each a
    dict {"pos": [N,3], "vel": [N,3]} of CUDA tensors.  Returns the reference's 13 outputs by name."""
    left, center, right = frames
    m = center["pos"].mean(0, keepdim=True)                     # normalize_point_cloud: centroid-centred, scale 1
    h = 1.0
    pos_c, pos_l, pos_r = center["pos"] - m, left["pos"] - m, right["pos"] - m
    patch_idx, fps_idx = sample_patch_with_fps(pos_c, sample_num, seed_idx, fps_start)
    hi = {k: v[patch_idx] for k, v in (("pos_left", pos_l), ("pos", pos_c), ("pos_right", pos_r), ("vel_left", left["vel"]),
                                       ("vel", center["vel"]), ("vel_right", right["vel"]))}
    n_lo = fps_idx.shape[0]

    def noise():
        return torch.randn((n_lo, 3), device=pos_c.device, generator=generator) * jitter

    out = {"highres_pos_left": hi["pos_left"], "highres_pos": hi["pos"], "highres_pos_right": hi["pos_right"],
           "highres_vel_left": hi["vel_left"], "highres_vel": hi["vel"], "highres_vel_right": hi["vel_right"],
           "lowres_pos": hi["pos"][fps_idx] + noise(), "lowres_pos_left": hi["pos_left"][fps_idx] + noise(),
           "lowres_pos_right": hi["pos_right"][fps_idx] + noise(),
           # the reference indexes the un-patched velocity arrays with fps_idx here (tempo_dataset.py:98-100)
           "lowres_vel": center["vel"][fps_idx], "lowres_vel_left": left["vel"][fps_idx], "lowres_vel_right": right["vel"][fps_idx],
           "h": h, "patch_idx": patch_idx, "fps_idx": fps_idx}
    return out


def load_frame_npz(path: str, device="cuda") -> Dict[str, torch.Tensor]:
    """One simulation frame in the reference's npz format {pos, vel} (process_training_data.py:75-79)."""
    d = np.load(path)
    return {"pos": torch.from_numpy(d["pos"].astype(np.float32)).to(device), "vel": torch.from_numpy(d["vel"].astype(np.float32)).to(device)}
