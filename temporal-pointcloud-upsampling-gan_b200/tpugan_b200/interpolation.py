"""GPU replacement for ``gcn_lib.cubic_interpolation`` (gcn_lib/interpolation.py:103-123).

Same signature as the reference for one sample, plus a batched form that replaces the
frames x samples Python loop of ``interpolate_vel_lst`` (train_step_final.py:51-66).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import functional as F

DT = 0.025  # train_step_final.py:7


def cubic_interpolation(query_pos: torch.Tensor, field: torch.Tensor, pos: torch.Tensor, cutoff: float) -> torch.Tensor:
    """query_pos [Q,3], field [P,F], pos [P,3] -> [Q,F]; no gradient (the reference calls it
    under torch.no_grad())."""
    with torch.no_grad():
        out = F.cubic_interp(query_pos.detach().contiguous().float()[None], field.detach().contiguous().float()[None],
                             pos.detach().contiguous().float()[None], float(cutoff))
    return out[0]


def cubic_interpolation_batched(query_pos: torch.Tensor, field: torch.Tensor, pos: torch.Tensor,
                                cutoff: float) -> torch.Tensor:
    """query_pos [S,Q,3], field [S,P,F], pos [S,P,3] -> [S,Q,F] in one launch sequence."""
    with torch.no_grad():
        return F.cubic_interp(query_pos.detach().contiguous().float(), field.detach().contiguous().float(),
                              pos.detach().contiguous().float(), float(cutoff))


def interpolate_vel_lst(pred_pos_lst, gt_pos_lst, gt_vel_lst, opt, furthest_distance) -> Tuple[List, List]:
    """Drop-in for train_step_final.interpolate_vel_lst (:51-66): one batched call per frame."""
    gt_adv_lst, pred_adv_lst = [], []
    cutoff = 1.6 * opt.R / furthest_distance
    for f in range(len(pred_pos_lst)):
        with torch.no_grad():
            highres_adv = gt_vel_lst[f] * DT
            pred_adv = cubic_interpolation_batched(pred_pos_lst[f], highres_adv, gt_pos_lst[f], cutoff)
        gt_adv_lst.append(highres_adv)
        pred_adv_lst.append(pred_adv)
    return gt_adv_lst, pred_adv_lst
