"""torch(CPU) <-> numpy oracle glue with autograd where the reference back-propagates."""
import numpy as np
import torch

import oracle
from oracle.shims import recorder


def _np(t):
    return t.detach().cpu().numpy()


def _c(a):
    """copy for the record unless only shapes are kept"""
    return a if recorder.shapes_only or not recorder.enabled else a


def knn(p1, p2, K, lengths1=None, lengths2=None):
    d, i = oracle.knn(_np(p1), _np(p2), K, None if lengths1 is None else _np(lengths1),
                      None if lengths2 is None else _np(lengths2))
    recorder.record("knn", dict(p1=_c(_np(p1)), p2=_c(_np(p2)), K=K), dict(dists=d, idx=i))
    return torch.from_numpy(d), torch.from_numpy(i)


def frnn(p1, p2, K, r):
    rr = _np(r) if isinstance(r, torch.Tensor) else r
    d, i = oracle.frnn(_np(p1), _np(p2), K, rr)
    recorder.record("frnn", dict(p1=_c(_np(p1)), p2=_c(_np(p2)), K=K, r=np.float32(rr)), dict(dists=d, idx=i))
    return torch.from_numpy(d), torch.from_numpy(i)


def ball_query(radius, nsample, xyz, new_xyz):
    o = oracle.ball_query(radius, nsample, _np(xyz), _np(new_xyz))
    recorder.record("ball_query", dict(xyz=_c(_np(xyz)), new_xyz=_c(_np(new_xyz)), radius=np.float32(radius),
                                       nsample=nsample), dict(idx=o))
    return torch.from_numpy(o)


def fps(xyz, npoint):
    o = oracle.fps(_np(xyz), npoint)
    recorder.record("fps", dict(xyz=_c(_np(xyz)), npoint=npoint), dict(idx=o))
    return torch.from_numpy(o)


class Grouping(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.N = features.shape[2]
        ctx.save_for_backward(idx)
        ctx.call_id = recorder.new_id()
        out = oracle.group_fwd(_np(features), _np(idx))
        recorder.record("group", dict(f=_c(_np(features)), idx=_c(_np(idx)), id=ctx.call_id), dict(out=out))
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = oracle.group_bwd(_np(grad_out.contiguous()), _np(idx), ctx.N)
        recorder.record("group_bwd", dict(grad_out=_c(_np(grad_out)), idx=_c(_np(idx)), N=ctx.N, fwd_id=ctx.call_id), dict(grad_f=g))
        return torch.from_numpy(g), None


class Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.N = features.shape[2]
        ctx.save_for_backward(idx)
        ctx.call_id = recorder.new_id()
        out = oracle.group_fwd(_np(features), _np(idx)[:, :, None])[..., 0]
        recorder.record("gather", dict(f=_c(_np(features)), idx=_c(_np(idx)), id=ctx.call_id), dict(out=out))
        return torch.from_numpy(np.ascontiguousarray(out))

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = oracle.group_bwd(_np(grad_out.contiguous())[..., None], _np(idx)[:, :, None], ctx.N)
        recorder.record("gather_bwd", dict(grad_out=_c(_np(grad_out)), idx=_c(_np(idx)), N=ctx.N, fwd_id=ctx.call_id),
                        dict(grad_f=g))
        return torch.from_numpy(g), None


class ChamferSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, tgt, directions):
        r = oracle.chamfer_fwd(_np(src), _np(tgt), directions)
        ctx.directions = directions
        ctx.save_for_backward(src, tgt, torch.from_numpy(r["i_src"]), torch.from_numpy(r["i_tgt"]))
        recorder.record("chamfer", dict(src=_c(_np(src)), tgt=_c(_np(tgt)), directions=directions),
                        dict(sum_src=r["sum_src"], sum_tgt=r["sum_tgt"], i_src=r["i_src"], i_tgt=r["i_tgt"]))
        return torch.from_numpy(r["sum_src"]), torch.from_numpy(r["sum_tgt"])

    @staticmethod
    def backward(ctx, g_src, g_tgt):
        src, tgt, i_s, i_t = ctx.saved_tensors
        gs, gt = oracle.chamfer_bwd(_np(src), _np(tgt), _np(i_s), _np(i_t), _np(g_src), _np(g_tgt), ctx.directions)
        recorder.record("chamfer_bwd", dict(src=_c(_np(src)), tgt=_c(_np(tgt)), i_src=_c(_np(i_s)), i_tgt=_c(_np(i_t)),
                                            g_src=_c(_np(g_src)), g_tgt=_c(_np(g_tgt)),
                                            directions=ctx.directions), dict(grad_src=gs, grad_tgt=gt))
        return torch.from_numpy(gs), torch.from_numpy(gt), None
