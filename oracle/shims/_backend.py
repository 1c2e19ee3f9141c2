"""torch(CPU) <-> numpy oracle glue with autograd where the reference back-propagates."""
import numpy as np
import torch

import oracle
from oracle.shims import recorder


def _np(t):
    return t.detach().cpu().numpy()


def _sha(t):
    import hashlib

    return hashlib.sha1(np.ascontiguousarray(_np(t)).tobytes()).hexdigest()[:16] if recorder.enabled else None


def _out(*tensors):
    """tag freshly made output tensors with the call recorded last (dependency tracking)"""
    recorder.tag(*tensors)
    return tensors[0] if len(tensors) == 1 else tensors


def _c(a):
    """copy for the record unless only shapes are kept"""
    return a if recorder.shapes_only or not recorder.enabled else a


# "c" = the C oracle (default); "torch" = the pure-torch dense formulations (oracle/torch_formulations.py) for the
# ops the generator forward uses (kNN, grouping) -- the BASELINE configs[0] CPU leg of bench.py
IMPL = {"mode": "c"}


def knn(p1, p2, K, lengths1=None, lengths2=None):
    if IMPL["mode"] == "torch" and lengths1 is None and lengths2 is None:
        from oracle import torch_formulations as tf

        return tf.knn(p1, p2, K)
    deps = recorder.deps(p1=p1, p2=p2)
    d, i = oracle.knn(_np(p1), _np(p2), K, None if lengths1 is None else _np(lengths1),
                      None if lengths2 is None else _np(lengths2))
    # content digests: calls that received bit-identical clouds (IDGCNLayer runs three searches on one feature
    # map, gcn.py:258-265) are replayed on bit-identical inputs
    recorder.record("knn", dict(p1=_c(_np(p1)), p2=_c(_np(p2)), K=K, deps=deps, p1_sha=_sha(p1), p2_sha=_sha(p2)),
                    dict(dists=d, idx=i))
    return _out(torch.from_numpy(d), torch.from_numpy(i))


def frnn(p1, p2, K, r):
    rr = _np(r) if isinstance(r, torch.Tensor) else r
    deps = recorder.deps(p1=p1, p2=p2)
    d, i = oracle.frnn(_np(p1), _np(p2), K, rr)
    recorder.record("frnn", dict(p1=_c(_np(p1)), p2=_c(_np(p2)), K=K, r=np.float32(rr), deps=deps), dict(dists=d, idx=i))
    return _out(torch.from_numpy(d), torch.from_numpy(i))


def ball_query(radius, nsample, xyz, new_xyz):
    deps = recorder.deps(xyz=xyz, new_xyz=new_xyz)
    o = oracle.ball_query(radius, nsample, _np(xyz), _np(new_xyz))
    recorder.record("ball_query", dict(xyz=_c(_np(xyz)), new_xyz=_c(_np(new_xyz)), radius=np.float32(radius),
                                       nsample=nsample, deps=deps), dict(idx=o))
    return _out(torch.from_numpy(o))


def fps(xyz, npoint):
    deps = recorder.deps(xyz=xyz)
    o = oracle.fps(_np(xyz), npoint)
    recorder.record("fps", dict(xyz=_c(_np(xyz)), npoint=npoint, deps=deps), dict(idx=o))
    return _out(torch.from_numpy(o))


class Grouping(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.N = features.shape[2]
        ctx.save_for_backward(idx)
        ctx.call_id = recorder.new_id()
        if IMPL["mode"] == "torch":
            from oracle import torch_formulations as tf

            return tf.grouping(features, idx)
        deps = recorder.deps(f=features, idx=idx)
        out = oracle.group_fwd(_np(features), _np(idx))
        recorder.record("group", dict(f=_c(_np(features)), idx=_c(_np(idx)), id=ctx.call_id, deps=deps), dict(out=out))
        return _out(torch.from_numpy(out))

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        deps = recorder.deps(grad_out=grad_out, idx=idx)
        g = oracle.group_bwd(_np(grad_out.contiguous()), _np(idx), ctx.N)
        recorder.record("group_bwd", dict(grad_out=_c(_np(grad_out)), idx=_c(_np(idx)), N=ctx.N, fwd_id=ctx.call_id, deps=deps),
                        dict(grad_f=g))
        return _out(torch.from_numpy(g)), None


class Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.N = features.shape[2]
        ctx.save_for_backward(idx)
        ctx.call_id = recorder.new_id()
        deps = recorder.deps(f=features, idx=idx)
        out = oracle.group_fwd(_np(features), _np(idx)[:, :, None])[..., 0]
        recorder.record("gather", dict(f=_c(_np(features)), idx=_c(_np(idx)), id=ctx.call_id, deps=deps), dict(out=out))
        return _out(torch.from_numpy(np.ascontiguousarray(out)))

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        deps = recorder.deps(grad_out=grad_out, idx=idx)
        g = oracle.group_bwd(_np(grad_out.contiguous())[..., None], _np(idx)[:, :, None], ctx.N)
        recorder.record("gather_bwd", dict(grad_out=_c(_np(grad_out)), idx=_c(_np(idx)), N=ctx.N, fwd_id=ctx.call_id, deps=deps),
                        dict(grad_f=g))
        return _out(torch.from_numpy(g)), None


class ChamferSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, tgt, directions):
        deps = recorder.deps(src=src, tgt=tgt)
        r = oracle.chamfer_fwd(_np(src), _np(tgt), directions)
        ctx.directions = directions
        i_s, i_t = torch.from_numpy(r["i_src"]), torch.from_numpy(r["i_tgt"])
        ctx.save_for_backward(src, tgt, i_s, i_t)
        recorder.record("chamfer", dict(src=_c(_np(src)), tgt=_c(_np(tgt)), directions=directions, deps=deps),
                        dict(sum_src=r["sum_src"], sum_tgt=r["sum_tgt"], i_src=r["i_src"], i_tgt=r["i_tgt"]))
        recorder.tag(i_s, i_t)
        return _out(torch.from_numpy(r["sum_src"]), torch.from_numpy(r["sum_tgt"]))

    @staticmethod
    def backward(ctx, g_src, g_tgt):
        src, tgt, i_s, i_t = ctx.saved_tensors
        deps = recorder.deps(src=src, tgt=tgt, i_src=i_s, g_src=g_src, g_tgt=g_tgt)
        gs, gt = oracle.chamfer_bwd(_np(src), _np(tgt), _np(i_s), _np(i_t), _np(g_src), _np(g_tgt), ctx.directions)
        recorder.record("chamfer_bwd", dict(src=_c(_np(src)), tgt=_c(_np(tgt)), i_src=_c(_np(i_s)), i_tgt=_c(_np(i_t)),
                                            g_src=_c(_np(g_src)), g_tgt=_c(_np(g_tgt)),
                                            directions=ctx.directions, deps=deps), dict(grad_src=gs, grad_tgt=gt))
        return _out(torch.from_numpy(gs), torch.from_numpy(gt)) + (None,)
