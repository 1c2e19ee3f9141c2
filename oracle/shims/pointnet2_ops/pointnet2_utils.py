import torch
import torch.nn as nn

from oracle.shims import _backend as B


def furthest_point_sample(xyz, npoint):
    return B.fps(xyz.contiguous(), npoint)


gather_operation = B.Gather.apply
grouping_operation = B.Grouping.apply


def ball_query(radius, nsample, xyz, new_xyz):
    return B.ball_query(radius, nsample, xyz.contiguous(), new_xyz.contiguous())


class QueryAndGroup(nn.Module):
    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features.contiguous(), idx)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz


class GroupAll(nn.Module):
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz
