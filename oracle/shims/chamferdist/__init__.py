import torch

from oracle.shims import _backend as B


class ChamferDistance(torch.nn.Module):
    def forward(self, source_cloud, target_cloud, bidirectional=False, reverse=False, batch_reduction="mean",
                point_reduction="sum"):
        directions = 3 if bidirectional else (2 if reverse else 1)
        s, t = B.ChamferSums.apply(source_cloud.contiguous().float(), target_cloud.contiguous().float(), directions)

        def red(x, n):
            if point_reduction == "mean":
                x = x / n
            return x.mean() if batch_reduction == "mean" else (x.sum() if batch_reduction == "sum" else x)

        f, b = red(s, source_cloud.shape[1]), red(t, target_cloud.shape[1])
        return f + b if bidirectional else (b if reverse else f)
