from oracle.shims import _backend as B


def frnn_grid_points(points1, points2, lengths1=None, lengths2=None, K=-1, r=-1, grid=None, return_nn=True,
                     return_sorted=True, radius_cell_ratio=2.0):
    d, i = B.frnn(points1.contiguous(), points2.contiguous(), K, r)
    return d, i, None, None
