"""Import-only stand-in for open3d (CPU data-prep / visualisation in train_utils.py)."""


class _Unavailable:
    def __getattr__(self, name):
        raise NotImplementedError("open3d is not installed (tpugan_b200 import stub)")


geometry = utility = visualization = io = _Unavailable()
