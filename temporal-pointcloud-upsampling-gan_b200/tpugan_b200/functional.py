"""Tensor-level front-end of the sm_100a kernels (PyTorch is plumbing: memory + streams).

Every function validates its arguments the way the upstream extension it replaces
does, allocates outputs/workspace with the caching allocator, and launches on
``torch.cuda.current_stream()``.  No function synchronises or reads back.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import _lib
from .recording import log as _log

_vp = ctypes.c_void_p


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else _vp(t.data_ptr())


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _req(t, name, dtype, ndim=None, exc=RuntimeError):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise exc(f"{name} must be a CUDA tensor (tpugan_b200 has no CPU path)")
    if t.dtype != dtype:
        raise exc(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise exc(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise exc(f"{name} must have {ndim} dimensions, got shape {tuple(t.shape)}")


class _on_device:
    """`with torch.cuda.device(d)` costs several microseconds per call; the common case (tensor on the
    current device) needs no guard at all."""
    __slots__ = ("guard",)

    def __init__(self, device):
        self.guard = None if device.index is None or device.index == torch.cuda.current_device() \
            else torch.cuda.device(device)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _ws(nbytes: int, device) -> Optional[torch.Tensor]:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def _lengths(lengths, B, P, device, name):
    if lengths is None:
        return None
    if not isinstance(lengths, torch.Tensor):
        lengths = torch.as_tensor(lengths)
    if lengths.shape != (B,):
        raise ValueError(f"{name} must have shape (N,).")
    return lengths.to(device=device, dtype=torch.int64).contiguous()


# --------------------------------------------------------------------------- kNN / FRNN
def knn(p1, p2, K: int, lengths1=None, lengths2=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K nearest by squared L2, (d2, idx) ascending.  p1 [B,P1,D], p2 [B,P2,D] ->
    dists [B,P1,K] f32, idx [B,P1,K] int64.  (pytorch3d knn_points; gcn_lib/pointnet/gcn.py:16)"""
    _req(p1, "p1", torch.float32, 3)
    _req(p2, "p2", torch.float32, 3)
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    l1 = _lengths(lengths1, B, P1, p1.device, "lengths1")
    l2 = _lengths(lengths2, B, P2, p1.device, "lengths2")
    rec = _log.begin("knn", p1=p1, p2=p2, K=int(K), lengths1=l1, lengths2=l2) if _log.on else None
    if knn_memo.enabled and l1 is None and l2 is None and knn_memo.eligible(p1, p2, K):
        dists, idx = knn_memo.run(p1, p2, K)
    else:
        dists, idx = _knn_raw(p1, p2, K, l1, l2)
    if rec is not None:
        _log.end(rec, dists=dists, idx=idx)
    return dists, idx


def _knn_raw(p1, p2, K, l1=None, l2=None, skip_flag=None):
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    dists = torch.empty((B, P1, K), dtype=torch.float32, device=p1.device)
    idx = torch.empty((B, P1, K), dtype=torch.int64, device=p1.device)
    with _on_device(p1.device):
        nbytes = _lib.load().tpg_knn_workspace_bytes(B, P1, P2, D, K)  # > 0: tensor-core (tcgen05) path
        ws = _ws(nbytes, p1.device) if nbytes else None
        _lib.call("tpg_knn_cond_f32", _ptr(p1), _ptr(p2), _ptr(l1), _ptr(l2), B, P1, P2, D, K, _ptr(dists), _ptr(idx),
                  _ptr(ws), nbytes, _ptr(skip_flag), _stream())
    return dists, idx


class _KnnMemo:
    """Exact memoisation of repeated feature-space searches inside one train step (opt-in).

    The reference's IDGCNLayer runs knn_points three times on bit-identical feature maps (K = 9, 20, 20;
    gcn_lib/pointnet/gcn.py:258-265), each on a fresh ``.contiguous()`` copy.  With ``enabled`` a call on the
    tensor-core path is compared on the device (bytes, not pointers) with the inputs of the most recent call of the
    same shape; the search kernels of the new call return at once when the contents match, and the result is
    the prefix of the cached list (the K nearest are a prefix of the Kc >= K nearest in the canonical order).
    Searches with K < 20 are computed with 20 neighbours so that the following K = 20 calls can reuse them.
    No host synchronisation, capturable; ``clear()`` at the end of a step (nothing survives a step)."""

    KMIN = 20

    def __init__(self, capacity: int = 4):
        self.enabled = False
        self.capacity = capacity
        self.entries = []  # (p1, p2, v1, v2, Kc, dists, idx, event)
        self._retired = []  # evicted entries stay alive until clear(): another stream may still read them

    @staticmethod
    def eligible(p1, p2, K):
        D = p1.shape[2]
        return D in (32, 64) and p2.shape[1] >= 1024 and K <= 24 and p1.is_contiguous() and p2.is_contiguous() \
            and p1.data_ptr() % 16 == 0 and p2.data_ptr() % 16 == 0 \
            and p1.shape[0] * p1.shape[1] * ((p2.shape[1] + 127) // 128) * 16 <= (1 << 30)  # knn_feat_eligible

    def clear(self):
        self.entries.clear()
        self._retired.clear()

    def _remember(self, p1, p2, Kc, dists, idx):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(p1.device))
        self.entries.append((p1, p2, p1._version, p2._version, Kc, dists, idx, ev))
        self._retired.extend(self.entries[:-self.capacity])
        del self.entries[:-self.capacity]
        # evicted entries only have to outlive the kernels that may still read them (another stream may be
        # comparing against them): drop those whose event has completed, so that an uncaptured train loop over
        # the drop-in packages does not accumulate every call's tensors.  Under stream capture events cannot be
        # queried; a captured step calls clear() when it ends.
        if self._retired and not torch.cuda.is_current_stream_capturing():
            self._retired = [e for e in self._retired if not e[7].query()]

    def run(self, p1, p2, K):
        dev = p1.device
        Kc = max(K, self.KMIN)
        hit = None
        for e in reversed(self.entries):  # the most recent call of this shape (a layer's searches are consecutive)
            if e[0].shape == p1.shape and e[1].shape == p2.shape and e[0].device == dev:
                if e[4] >= Kc and e[0]._version == e[2] and e[1]._version == e[3]:
                    hit = e
                break
        if hit is None:
            dists, idx = _knn_raw(p1, p2, Kc)
        else:
            cp1, cp2, _, _, Kh, cd, ci, ev = hit
            torch.cuda.current_stream(dev).wait_event(ev)
            flag = torch.ones(1, dtype=torch.int32, device=dev)
            with _on_device(dev):
                _lib.call("tpg_bytes_equal_and", _ptr(p1), _ptr(cp1), p1.numel() * 4, _ptr(flag), _stream())
                if not (p2 is p1 and cp2 is cp1):
                    _lib.call("tpg_bytes_equal_and", _ptr(p2), _ptr(cp2), p2.numel() * 4, _ptr(flag), _stream())
            dists, idx = _knn_raw(p1, p2, Kc, skip_flag=flag)
            with _on_device(dev):
                _lib.call("tpg_knn_take_prefix", _ptr(flag), _ptr(cd), _ptr(ci), Kh, _ptr(dists), _ptr(idx), Kc,
                          p1.shape[0] * p1.shape[1], _stream())
        self._remember(p1, p2, Kc, dists, idx)
        # the caller always gets private copies: the reference edits neighbour lists in place
        # (gcn_lib/pointnet/gcn.py:44, discriminator.py:39), which must never reach the cached result
        if Kc == K:
            return dists.clone(), idx.clone()
        return dists[:, :, :K].contiguous(), idx[:, :, :K].contiguous()


knn_memo = _KnnMemo()


def frnn(p1, p2, K: int, r, lengths1=None, lengths2=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K nearest with d2 < r^2, -1 padded.  (frnn.frnn_grid_points; discriminator.py:27)"""
    _req(p1, "points1", torch.float32, 3, exc=TypeError)
    _req(p2, "points2", torch.float32, 3, exc=TypeError)
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    B, P1, D = p1.shape
    P2 = p2.shape[1]
    l1 = _lengths(lengths1, B, P1, p1.device, "lengths1")
    l2 = _lengths(lengths2, B, P2, p1.device, "lengths2")
    r_dev = None
    r_host = 0.0
    if isinstance(r, torch.Tensor):
        if r.numel() == 1 and not r.is_cuda:
            r_host = float(r)
        else:
            r_dev = r.to(device=p1.device, dtype=torch.float32).expand(B).contiguous()
    else:
        r_host = float(r)
    rec = _log.begin("frnn", p1=p1, p2=p2, K=int(K), r=(r_dev if r_dev is not None else r_host), lengths1=l1,
                     lengths2=l2) if _log.on else None
    dists = torch.empty((B, P1, K), dtype=torch.float32, device=p1.device)
    idx = torch.empty((B, P1, K), dtype=torch.int64, device=p1.device)
    with _on_device(p1.device):
        nbytes = _lib.load().tpg_frnn_workspace_bytes(B, P1, P2, D, K)
        ws = _ws(nbytes, p1.device)
        _lib.call("tpg_frnn_f32", _ptr(p1), _ptr(p2), _ptr(l1), _ptr(l2), B, P1, P2, D, K, r_host, _ptr(r_dev),
                  _ptr(dists), _ptr(idx), _ptr(ws), ws.numel(), _stream())
    if rec is not None:
        _log.end(rec, dists=dists, idx=idx)
    return dists, idx


def ball_query(radius: float, nsample: int, xyz, new_xyz) -> torch.Tensor:
    """pointnet2 ball_query: xyz [B,N,3], new_xyz [B,M,3] -> int32 [B,M,nsample]."""
    _req(xyz, "xyz", torch.float32, 3)
    _req(new_xyz, "new_xyz", torch.float32, 3)
    if xyz.shape[2] != 3 or new_xyz.shape[2] != 3 or xyz.shape[0] != new_xyz.shape[0]:
        raise RuntimeError("ball_query expects xyz (B,N,3) and new_xyz (B,M,3)")
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    rec = _log.begin("ball_query", xyz=xyz, new_xyz=new_xyz, radius=float(radius), nsample=int(nsample)) \
        if _log.on else None
    idx = torch.empty((B, M, nsample), dtype=torch.int32, device=xyz.device)
    with _on_device(xyz.device):
        nbytes = _lib.load().tpg_ball_query_workspace_bytes(B, N, M, int(nsample))  # > 0: uniform-grid search
        ws = _ws(nbytes, xyz.device) if nbytes else None
        _lib.call("tpg_ball_query_f32", _ptr(xyz), _ptr(new_xyz), B, N, M, float(radius), int(nsample), _ptr(idx),
                  _ptr(ws), nbytes, _stream())
    if rec is not None:
        _log.end(rec, idx=idx)
    return idx


# --------------------------------------------------------------------------- FPS
def fps(xyz, npoint: int) -> torch.Tensor:
    """pointnet2 furthest_point_sample: xyz [B,N,3] -> int32 [B,npoint] (discriminator.py:114)."""
    _req(xyz, "xyz", torch.float32, 3)
    if xyz.shape[2] != 3:
        raise RuntimeError("furthest_point_sample expects xyz (B,N,3)")
    B, N, _ = xyz.shape
    rec = _log.begin("fps", xyz=xyz, npoint=int(npoint)) if _log.on else None
    out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    with _on_device(xyz.device):
        nbytes = _lib.load().tpg_fps_workspace_bytes(B, N)
        ws = _ws(nbytes, xyz.device)
        _lib.call("tpg_fps_f32", _ptr(xyz), B, N, int(npoint), _ptr(out), _ptr(ws), ws.numel() if nbytes else 0,
                  _stream())
    if rec is not None:
        _log.end(rec, idx=out)
    return out


def fps_start(pts, k: int, start, return_rows: bool = False):
    """sampling.py mode: pts [B,N,D<=3], start [B] -> int64 [B,k] (+ rows [B,k,N])."""
    _req(pts, "pts", torch.float32, 3)
    B, N, D = pts.shape
    start = torch.as_tensor(start, dtype=torch.int64, device=pts.device).expand(B).contiguous()
    rec = _log.begin("fps_start", pts=pts, k=int(k), start=start) if _log.on else None
    out = torch.empty((B, k), dtype=torch.int64, device=pts.device)
    rows = torch.empty((B, k, N), dtype=torch.float32, device=pts.device) if return_rows else None
    with _on_device(pts.device):
        nbytes = _lib.load().tpg_fps_workspace_bytes(B, N)
        ws = _ws(nbytes, pts.device)
        _lib.call("tpg_fps_start_f32", _ptr(pts), B, N, D, int(k), _ptr(start), _ptr(out), _ptr(rows), _ptr(ws),
                  ws.numel() if nbytes else 0, _stream())
    if rec is not None:
        _log.end(rec, idx=out, rows=rows)
    return (out, rows) if return_rows else out


# --------------------------------------------------------------------------- grouping
def group_fwd(f, idx, center=None, _op: str = "group") -> torch.Tensor:
    """f [B,C,N], idx int32 [B,M,k] -> [B,C,M,k] (optionally minus center [B,C,M])."""
    _req(f, "features", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    B, C, N = f.shape
    if idx.shape[0] != B:
        raise RuntimeError("grouping_operation: batch mismatch between features and idx")
    _, M, k = idx.shape
    if center is not None:
        _req(center, "center", torch.float32, 3)
        if tuple(center.shape) != (B, C, M):
            raise RuntimeError("group_fwd: center must be [B,C,M]")
    rec = _log.begin(_op, f=f, idx=idx, center=center) if _log.on else None
    out = torch.empty((B, C, M, k), dtype=torch.float32, device=f.device)
    with _on_device(f.device):
        nbytes = _lib.load().tpg_group_fwd_workspace_bytes(B, C, N, M, k)  # > 0: rows too long for shared memory
        if nbytes:
            ws = _ws(nbytes, f.device)
            _lib.call("tpg_group_fwd_ws_f32", _ptr(f), _ptr(idx), _ptr(center), B, C, N, M, k, _ptr(out), _ptr(ws), nbytes,
                      _stream())
        else:
            _lib.call("tpg_group_fwd_f32", _ptr(f), _ptr(idx), _ptr(center), B, C, N, M, k, _ptr(out), _stream())
    if rec is not None:
        _log.end(rec, out=out)
    return out


def inverse_index(idx, N: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR of idx int32 [B, ...] with keys in [0,N): (seg_offsets [B,N+1], seg_items [B,L])."""
    _req(idx, "idx", torch.int32)
    B = idx.shape[0]
    L = idx.numel() // max(B, 1)
    off = torch.empty((B, N + 1), dtype=torch.int32, device=idx.device)
    items = torch.empty((B, max(L, 1)), dtype=torch.int32, device=idx.device)
    with _on_device(idx.device):
        nbytes = _lib.load().tpg_inverse_index_workspace_bytes(B, N, L)
        ws = _ws(nbytes, idx.device)
        _lib.call("tpg_inverse_index_build", _ptr(idx), B, N, L, _ptr(off), _ptr(items), _ptr(ws), ws.numel(),
                  _stream())
    return off, items


class _CsrCache:
    """Inverse indices keyed by the idx tensor they were built from.  Holding the idx
    tensor keeps its storage alive, so (data_ptr, version) cannot alias a new tensor.

    With ``prefetch_enabled`` the grouping / gather *forward* already builds the inverse index its backward
    will need, on a side stream that only waits for idx: the build leaves the critical path of the backward
    pass (a chain of CSR build -> gradient per layer) and fills idle SMs during the forward.  ``get`` makes
    the consuming stream wait for the build; ``join`` (end of a captured step) waits for builds nobody used."""

    def __init__(self, capacity: int = 16):
        self.capacity = capacity
        self.entries: "OrderedDict[tuple, tuple]" = OrderedDict()
        self.prefetch_enabled = False
        self._pools = {}
        self._rr = 0

    @staticmethod
    def _key(idx, N, segments=1):
        return (idx.data_ptr(), idx._version, tuple(idx.shape), N, idx.device.index, segments)

    def _insert(self, key, entry):
        self.entries[key] = entry
        cap = max(self.capacity, 256) if self.prefetch_enabled else self.capacity
        while len(self.entries) > cap:
            self.entries.popitem(last=False)

    def prefetch(self, idx: torch.Tensor, N: int) -> None:
        key = self._key(idx, N)
        if key in self.entries:
            return
        dev = idx.device
        pool = self._pools.get(dev.index)
        if pool is None:
            pool = self._pools[dev.index] = [torch.cuda.Stream(device=dev) for _ in range(4)]
        side = pool[self._rr % len(pool)]
        self._rr += 1
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        side.wait_event(ready)
        with torch.cuda.stream(side):
            off, items = inverse_index(idx, N)
            done = torch.cuda.Event()
            done.record(side)
        self._insert(key, (idx, off, items, done))

    def get(self, idx: torch.Tensor, N: int, segments: int = 1):
        """(seg_offsets, seg_items) of idx; segments > 1: the inverse index of idx viewed as [B*segments, L/segments]
        (layout of tpg_group_bwd_segmented_f32)."""
        key = self._key(idx, N, segments)
        hit = self.entries.get(key)
        if hit is not None:
            self.entries.move_to_end(key)
            if hit[3] is not None:
                torch.cuda.current_stream(idx.device).wait_event(hit[3])
            return hit[1], hit[2]
        off, items = inverse_index(idx if segments == 1 else idx.reshape(idx.shape[0] * segments, -1), N)
        self._insert(key, (idx, off, items, None))
        return off, items

    def join(self) -> None:
        for e in self.entries.values():
            if e[3] is not None:
                torch.cuda.current_stream(e[0].device).wait_event(e[3])

    def clear(self):
        self.entries.clear()


csr_cache = _CsrCache()


def group_bwd(grad_out, off, items, N: int) -> torch.Tensor:
    """grad_out [B,C,M,k] (or [B,C,L]) -> grad_f [B,C,N] through the CSR."""
    _req(grad_out, "grad_out", torch.float32)
    B, C = grad_out.shape[0], grad_out.shape[1]
    L = grad_out.numel() // max(B * C, 1)
    gf = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out.device):
        _lib.call("tpg_group_bwd_f32", _ptr(grad_out), _ptr(off), _ptr(items), B, C, N, L, _ptr(gf), _stream())
    return gf


def group_bwd_auto(grad_out, idx, N: int) -> torch.Tensor:
    """Grouping backward from the index tensor: grad_out [B,C,M,k], idx int32 [B,M,k] -> grad_f [B,C,N].  Picks the
    inverse-index layout the kernels want (whole rows, or S segments for rows of more than 65536 positions) and
    shares it through csr_cache."""
    B, C = grad_out.shape[0], grad_out.shape[1]
    L = grad_out.numel() // max(B * C, 1)
    S = _lib.load().tpg_group_bwd_segments(B, C, N, L) if grad_out.is_cuda else 0
    if S <= 1:
        off, items = csr_cache.get(idx, N)
        return group_bwd(grad_out, off, items, N)
    off, items = csr_cache.get(idx, N, segments=S)
    _req(grad_out, "grad_out", torch.float32)
    gf = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out.device):
        _lib.call("tpg_group_bwd_segmented_f32", _ptr(grad_out), _ptr(off), _ptr(items), B, C, N, L, S, _ptr(gf), _stream())
    return gf


def group_reduce_fwd(f, idx, op: int = _lib.REDUCE_MAX, want_arg: bool = True):
    """Fused gather + reduce over k: f [B,C,N], idx int32 [B,M,k] -> out [B,C,M] (+ arg int32)."""
    _req(f, "features", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    B, C, N = f.shape
    _, M, k = idx.shape
    rec = _log.begin("group_reduce", f=f, idx=idx, op=int(op)) if _log.on else None
    out = torch.empty((B, C, M), dtype=torch.float32, device=f.device)
    arg = torch.empty((B, C, M), dtype=torch.int32, device=f.device) if (want_arg and op != _lib.REDUCE_SUM) else None
    with _on_device(f.device):
        nbytes = _lib.load().tpg_group_reduce_workspace_bytes(B, C, N)  # > 0: rows too long for shared memory
        ws = _ws(nbytes, f.device) if nbytes else None
        _lib.call("tpg_group_reduce_fwd_ws_f32", _ptr(f), _ptr(idx), None, B, C, N, M, k, int(op), _ptr(out), _ptr(arg),
                  _ptr(ws), nbytes, _stream())
    if rec is not None:
        _log.end(rec, out=out, arg=arg)
    return out, arg


def group_reduce_bwd(grad_out, arg, off, items, N: int, k: int, op: int) -> torch.Tensor:
    _req(grad_out, "grad_out", torch.float32, 3)
    B, C, M = grad_out.shape
    gf = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out.device):
        _lib.call("tpg_group_reduce_bwd_f32", _ptr(grad_out), _ptr(arg), _ptr(off), _ptr(items), B, C, N, M, k,
                  int(op), _ptr(gf), _stream())
    return gf


# --------------------------------------------------------------------------- conv-input assembly (K11 / K12)
def group_assemble(parts, idx) -> torch.Tensor:
    """One-pass assembly of a concatenated conv input (include/tpugan_b200.h: tpg_group_assemble_f32).
    parts: sequence of ("gather", src [B,C,N], center [B,C,M] or None) / ("broadcast", src [B,C,M], None);
    idx int32 [B,M,k] -> [B, sum C, M, k]."""
    _req(idx, "idx", torch.int32, 3)
    B, M, k = idx.shape
    if not 1 <= len(parts) <= _lib.ASSEMBLE_MAX_PARTS:
        raise RuntimeError(f"group_assemble: 1..{_lib.ASSEMBLE_MAX_PARTS} parts")
    arr = (_lib.AssemblePart * len(parts))()
    ctot = 0
    for n, (mode, src, center) in enumerate(parts):
        _req(src, f"part {n} source", torch.float32, 3)
        if src.shape[0] != B or src.device != idx.device:
            raise RuntimeError("group_assemble: batch / device mismatch between a part and idx")
        C = src.shape[1]
        if mode == "gather":
            if center is not None:
                _req(center, f"part {n} center", torch.float32, 3)
                if tuple(center.shape) != (B, C, M):
                    raise RuntimeError("group_assemble: center must be [B,C,M]")
            arr[n] = _lib.AssemblePart(src.data_ptr(), center.data_ptr() if center is not None else None, C,
                                       src.shape[2], _lib.PART_GATHER)
        elif mode == "broadcast":
            if src.shape[2] != M or center is not None:
                raise RuntimeError("group_assemble: a broadcast part is [B,C,M] and takes no center")
            arr[n] = _lib.AssemblePart(src.data_ptr(), None, C, 0, _lib.PART_BROADCAST)
        else:
            raise RuntimeError(f"group_assemble: unknown part mode {mode!r}")
        ctot += C
    rec = _log.begin("group_assemble", idx=idx, modes=[p[0] for p in parts], srcs=[p[1] for p in parts],
                     centers=[p[2] for p in parts]) if _log.on else None
    out = torch.empty((B, ctot, M, k), dtype=torch.float32, device=idx.device)
    with _on_device(idx.device):
        _lib.call("tpg_group_assemble_f32", ctypes.cast(arr, _vp), len(parts), _ptr(idx), B, M, k, _ptr(out), _stream())
    if rec is not None:
        _log.end(rec, out=out)
    return out


def edge_affine_fwd(p, q, center, idx, slope: float) -> torch.Tensor:
    """out[b,c,m,j] = p[b,c,i] + LeakyReLU_slope(q[b,c,i] - center[b,c,m]), i = idx[b,m,j]  (tpg_edge_affine_fwd_f32)."""
    _req(p, "p", torch.float32, 3)
    _req(q, "q", torch.float32, 3)
    _req(center, "center", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    B, C, N = q.shape
    _, M, k = idx.shape
    if tuple(p.shape) != (B, C, N) or tuple(center.shape) != (B, C, M) or idx.shape[0] != B:
        raise RuntimeError("edge_affine: p, q must be [B,C,N], center [B,C,M], idx [B,M,k]")
    rec = _log.begin("edge_affine", p=p, q=q, center=center, idx=idx, slope=float(slope)) if _log.on else None
    out = torch.empty((B, C, M, k), dtype=torch.float32, device=q.device)
    with _on_device(q.device):
        _lib.call("tpg_edge_affine_fwd_f32", _ptr(p), _ptr(q), _ptr(center), _ptr(idx), float(slope), B, C, N, M, k,
                  _ptr(out), _stream())
    if rec is not None:
        _log.end(rec, out=out)
    return out


def edge_affine_bwd(grad_out, q, center, idx, slope: float, want_center: bool = True):
    """-> (g2 [B,C,M,k] = grad_out * LeakyReLU'(q[i] - center), grad_center [B,C,M] = -sum_j g2 or None)."""
    _req(grad_out, "grad_out", torch.float32, 4)
    B, C, N = q.shape
    _, M, k = idx.shape
    rec = _log.begin("edge_affine_bwd", grad_out=grad_out, q=q, center=center, idx=idx, slope=float(slope)) if _log.on else None
    g2 = torch.empty((B, C, M, k), dtype=torch.float32, device=q.device)
    gc = torch.empty((B, C, M), dtype=torch.float32, device=q.device) if want_center else None
    with _on_device(q.device):
        _lib.call("tpg_edge_affine_bwd_f32", _ptr(grad_out), _ptr(q), _ptr(center), _ptr(idx), float(slope), B, C, N, M, k,
                  _ptr(g2), _ptr(gc), _stream())
    if rec is not None:
        _log.end(rec, g2=g2, grad_center=gc)
    return g2, gc


# --------------------------------------------------------------------------- three_nn / interpolate
def three_nn(unknown, known):
    _req(unknown, "unknown", torch.float32, 3)
    _req(known, "known", torch.float32, 3)
    B, n, _ = unknown.shape
    m = known.shape[1]
    rec = _log.begin("three_nn", unknown=unknown, known=known) if _log.on else None
    dist = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    with _on_device(unknown.device):
        nbytes = _lib.load().tpg_three_nn_workspace_bytes(B, n, m)
        ws = _ws(nbytes, unknown.device) if nbytes else None
        _lib.call("tpg_three_nn_f32", _ptr(unknown), _ptr(known), B, n, m, _ptr(dist), _ptr(idx), _ptr(ws), nbytes,
                  _stream())
    if rec is not None:
        _log.end(rec, dist=dist, idx=idx)
    return dist, idx


def three_interpolate_fwd(f, idx, w):
    _req(f, "features", torch.float32, 3)
    _req(idx, "idx", torch.int32, 3)
    _req(w, "weight", torch.float32, 3)
    B, c, m = f.shape
    n = idx.shape[1]
    rec = _log.begin("three_interpolate", f=f, idx=idx, w=w) if _log.on else None
    out = torch.empty((B, c, n), dtype=torch.float32, device=f.device)
    with _on_device(f.device):
        nbytes = _lib.load().tpg_group_reduce_workspace_bytes(B, c, m)
        ws = _ws(nbytes, f.device) if nbytes else None
        _lib.call("tpg_group_reduce_fwd_ws_f32", _ptr(f), _ptr(idx), _ptr(w), B, c, m, n, 3, _lib.REDUCE_SUM, _ptr(out), None,
                  _ptr(ws), nbytes, _stream())
    if rec is not None:
        _log.end(rec, out=out)
    return out


def three_interpolate_bwd(grad_out, w, off, items, m: int):
    _req(grad_out, "grad_out", torch.float32, 3)
    B, c, n = grad_out.shape
    gf = torch.empty((B, c, m), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out.device):
        _lib.call("tpg_three_interpolate_bwd_f32", _ptr(grad_out), _ptr(w), _ptr(off), _ptr(items), B, c, m, n,
                  _ptr(gf), _stream())
    return gf


# --------------------------------------------------------------------------- Chamfer
def chamfer_fwd(src, tgt, directions: int, lengths_src=None, lengths_tgt=None):
    _req(src, "source_cloud", torch.float32, 3)
    _req(tgt, "target_cloud", torch.float32, 3)
    B, P1, D = src.shape
    P2 = tgt.shape[1]
    dev = src.device
    f, r = bool(directions & 1), bool(directions & 2)
    d_s = torch.empty((B, P1), dtype=torch.float32, device=dev) if f else None
    i_s = torch.empty((B, P1), dtype=torch.int32, device=dev) if f else None
    s_s = torch.empty((B,), dtype=torch.float32, device=dev) if f else None
    d_t = torch.empty((B, P2), dtype=torch.float32, device=dev) if r else None
    i_t = torch.empty((B, P2), dtype=torch.int32, device=dev) if r else None
    s_t = torch.empty((B,), dtype=torch.float32, device=dev) if r else None
    with _on_device(dev):
        nbytes = _lib.load().tpg_chamfer_fwd_workspace_bytes(B, P1, P2, D)  # > 0: uniform-grid search
        ws = _ws(nbytes, dev) if nbytes else None
        _lib.call("tpg_chamfer_fwd_f32", _ptr(src), _ptr(tgt), _ptr(lengths_src), _ptr(lengths_tgt), B, P1, P2, D,
                  int(directions), _ptr(d_s), _ptr(i_s), _ptr(d_t), _ptr(i_t), _ptr(s_s), _ptr(s_t), _ptr(ws), nbytes,
                  _stream())
    return dict(d_src=d_s, i_src=i_s, sum_src=s_s, d_tgt=d_t, i_tgt=i_t, sum_tgt=s_t)


def chamfer_bwd(src, tgt, i_src, i_tgt, g_src, g_tgt, directions: int, need_src=True, need_tgt=True,
                lengths_src=None, lengths_tgt=None):
    B, P1, D = src.shape
    P2 = tgt.shape[1]
    dev = src.device
    gs = torch.empty_like(src) if need_src else None
    gt = torch.empty_like(tgt) if need_tgt else None
    with _on_device(dev):
        nbytes = _lib.load().tpg_chamfer_bwd_workspace_bytes(B, P1, P2)
        ws = _ws(nbytes, dev)
        _lib.call("tpg_chamfer_bwd_f32", _ptr(src), _ptr(tgt), _ptr(lengths_src), _ptr(lengths_tgt), _ptr(i_src),
                  _ptr(i_tgt), _ptr(g_src), _ptr(g_tgt), B, P1, P2, D, int(directions), _ptr(gs), _ptr(gt), _ptr(ws),
                  ws.numel(), _stream())
    return gs, gt


# --------------------------------------------------------------------------- cubic interpolation
def cubic_interp(query, field, pos, cutoff: float) -> torch.Tensor:
    """Batched gcn_lib.cubic_interpolation: query [S,Q,3], field [S,P,F], pos [S,P,3] -> [S,Q,F]."""
    _req(query, "query_pos", torch.float32, 3)
    _req(field, "field", torch.float32, 3)
    _req(pos, "pos", torch.float32, 3)
    S, Q, _ = query.shape
    P, F = field.shape[1], field.shape[2]
    if pos.shape != (S, P, 3) or query.shape[2] != 3:
        raise ValueError("cubic_interp: expected query [S,Q,3], field [S,P,F], pos [S,P,3]")
    rec = _log.begin("cubic_interp", query=query, field=field, pos=pos, cutoff=float(cutoff)) if _log.on else None
    out = torch.empty((S, Q, F), dtype=torch.float32, device=query.device)
    with _on_device(query.device):
        nbytes = _lib.load().tpg_cubic_interp_workspace_bytes(S, Q, P)
        ws = _ws(nbytes, query.device)
        _lib.call("tpg_cubic_interp_f32", _ptr(query), _ptr(field), _ptr(pos), S, Q, P, F, float(cutoff), _ptr(out),
                  _ptr(ws), ws.numel(), _stream())
    if rec is not None:
        _log.end(rec, out=out)
    return out


def gather_rows(x, idx) -> torch.Tensor:
    """x [B,N,U], idx int64 [B,L] -> [B,L,U] (negative indices wrap)."""
    _req(x, "x", torch.float32, 3)
    _req(idx, "idx", torch.int64, 2)
    B, N, U = x.shape
    L = idx.shape[1]
    rec = _log.begin("gather_rows", x=x, idx=idx) if _log.on else None
    out = torch.empty((B, L, U), dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        _lib.call("tpg_gather_rows_f32", _ptr(x), _ptr(idx), B, N, U, L, _ptr(out), _stream())
    if rec is not None:
        _log.end(rec, out=out)
    return out


def _csr_of_rows(idx: torch.Tensor, N: int):
    """inverse index of an int64 row-index tensor idx [B,L] (negative = padding, clamped to key 0 and skipped by the
    consuming kernel)"""
    idx32 = idx.clamp_min(0).to(torch.int32).contiguous()
    return inverse_index(idx32, N)


def gather_rows_bwd(grad_out, idx, N: int) -> torch.Tensor:
    """backward of gather_rows: grad_out [B,L,U], idx int64 [B,L] -> grad_x [B,N,U] (ascending-l sums, no atomics)"""
    _req(grad_out, "grad_out", torch.float32, 3)
    _req(idx, "idx", torch.int64, 2)
    B, L, U = grad_out.shape
    rec = _log.begin("gather_rows_bwd", grad_out=grad_out, idx=idx, N=int(N)) if _log.on else None
    off, items = _csr_of_rows(idx, N)
    gx = torch.empty((B, N, U), dtype=torch.float32, device=grad_out.device)
    with _on_device(grad_out.device):
        _lib.call("tpg_gather_rows_bwd_f32", _ptr(grad_out), _ptr(idx), _ptr(off), _ptr(items), B, N, U, L, _ptr(gx),
                  _stream())
    if rec is not None:
        _log.end(rec, grad_x=gx)
    return gx


def knn_bwd(p1, p2, idx, grad_dists, lengths1=None, lengths2=None, need_p1=True, need_p2=True):
    """backward of the kNN / FRNN distances: (grad_p1 [B,P1,D] | None, grad_p2 [B,P2,D] | None)"""
    _req(p1, "p1", torch.float32, 3)
    _req(p2, "p2", torch.float32, 3)
    _req(idx, "idx", torch.int64, 3)
    _req(grad_dists, "grad_dists", torch.float32, 3)
    B, P1, D = p1.shape
    P2, K = p2.shape[1], idx.shape[2]
    rec = _log.begin("knn_bwd", p1=p1, p2=p2, idx=idx, grad_dists=grad_dists, lengths1=lengths1, lengths2=lengths2) \
        if _log.on else None
    gp1 = torch.empty_like(p1) if need_p1 else None
    gp2 = torch.empty_like(p2) if need_p2 else None
    off = items = None
    if need_p2:
        off, items = _csr_of_rows(idx.reshape(B, P1 * K), P2)
    with _on_device(p1.device):
        _lib.call("tpg_knn_bwd_f32", _ptr(p1), _ptr(p2), _ptr(idx), _ptr(grad_dists), _ptr(lengths1), _ptr(lengths2),
                  _ptr(off), _ptr(items), B, P1, P2, D, K, _ptr(gp1), _ptr(gp2), _stream())
    if rec is not None:
        _log.end(rec, grad_p1=gp1, grad_p2=gp2)
    return gp1, gp2


# =========================================================================== autograd
class NeighbourDists(torch.autograd.Function):
    """dists / idx of knn_points (r is None) or frnn_grid_points; gradient of dists w.r.t. both clouds through
    tpg_knn_bwd_f32.  Never taken on the reference's train step (no caller consumes `dists`: gcn.py:91,258;
    discriminator.py:33) but part of the drop-in surface."""

    @staticmethod
    def forward(ctx, p1, p2, lengths1, lengths2, K, r):
        dists, idx = knn(p1, p2, K, lengths1, lengths2) if r is None else frnn(p1, p2, K, r, lengths1, lengths2)
        B, P1, _ = p1.shape
        ctx.l1 = _lengths(lengths1, B, P1, p1.device, "lengths1")
        ctx.l2 = _lengths(lengths2, B, p2.shape[1], p1.device, "lengths2")
        ctx.save_for_backward(p1, p2, idx)
        ctx.mark_non_differentiable(idx)
        return dists, idx

    @staticmethod
    def backward(ctx, grad_dists, _grad_idx):
        p1, p2, idx = ctx.saved_tensors
        gp1, gp2 = knn_bwd(p1, p2, idx, grad_dists.contiguous(), ctx.l1, ctx.l2, ctx.needs_input_grad[0],
                           ctx.needs_input_grad[1])
        return gp1, gp2, None, None, None, None


class GatherRows(torch.autograd.Function):
    """x [B,N,U], idx int64 [B,L] (negative = padding -> zero row) -> [B,L,U]; knn_gather / frnn_gather."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.N = x.shape[1]
        ctx.save_for_backward(idx)
        ctx.has_pad = None
        return gather_rows(x, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return gather_rows_bwd(grad_out.contiguous(), idx, ctx.N), None


class GroupingOperation(torch.autograd.Function):
    """pointnet2_utils.GroupingOperation (gcn_lib/pointnet/gcn.py:207)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.N = features.shape[2]
        ctx.save_for_backward(idx)
        if csr_cache.prefetch_enabled and features.requires_grad:
            csr_cache.prefetch(idx, ctx.N)
        return group_fwd(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        rec = _log.begin("group_bwd", grad_out=grad_out, idx=idx, N=ctx.N) if _log.on else None
        gf = group_bwd_auto(grad_out, idx, ctx.N)
        if rec is not None:
            _log.end(rec, grad_f=gf)
        return gf, None


class GatherOperation(torch.autograd.Function):
    """pointnet2_utils.GatherOperation (discriminator.py:132): the k == 1 grouping."""

    @staticmethod
    def forward(ctx, features, idx):
        _req(idx, "idx", torch.int32, 2)
        ctx.N = features.shape[2]
        ctx.save_for_backward(idx)
        if csr_cache.prefetch_enabled and features.requires_grad:
            csr_cache.prefetch(idx, ctx.N)
        return group_fwd(features, idx.unsqueeze(-1), _op="gather").squeeze(-1)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        rec = _log.begin("gather_bwd", grad_out=grad_out, idx=idx, N=ctx.N) if _log.on else None
        off, items = csr_cache.get(idx, ctx.N)
        gf = group_bwd(grad_out, off, items, ctx.N)
        if rec is not None:
            _log.end(rec, grad_f=gf)
        return gf, None


class GroupReduce(torch.autograd.Function):
    """Fused grouping + max/sum/min over the neighbour axis (gcn_lib/pointnet/gcn.py:261-263)."""

    @staticmethod
    def forward(ctx, features, idx, op):
        out, arg = group_reduce_fwd(features, idx, op, want_arg=True)
        ctx.N, ctx.k, ctx.op = features.shape[2], idx.shape[2], op
        ctx.save_for_backward(idx, arg if arg is not None else idx)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, arg = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        arg = None if ctx.op == _lib.REDUCE_SUM else arg
        rec = _log.begin("group_reduce_bwd", grad_out=grad_out, idx=idx, arg=arg, N=ctx.N, op=ctx.op) if _log.on else None
        off, items = csr_cache.get(idx, ctx.N)
        gf = group_reduce_bwd(grad_out, arg, off, items, ctx.N, ctx.k, ctx.op)
        if rec is not None:
            _log.end(rec, grad_f=gf)
        return gf, None, None


class GroupAssemble(torch.autograd.Function):
    """cat of gathered (optionally centre-subtracted) and broadcast parts in one pass (K11):
    QueryAndGroup (discriminator.py:190) and FlowEmbedding's conv input (discriminator.py:270-277).
    apply(idx, modes, *tensors) with tensors = (src_0, center_0, src_1, center_1, ...), center None where unused."""

    @staticmethod
    def forward(ctx, idx, modes, *tensors):
        parts = [(m, tensors[2 * n], tensors[2 * n + 1]) for n, m in enumerate(modes)]
        ctx.modes = tuple(modes)
        ctx.shapes = [(p[1].shape[1], p[1].shape[2]) for p in parts]
        ctx.has_center = [p[2] is not None for p in parts]
        ctx.save_for_backward(idx)
        if csr_cache.prefetch_enabled:
            for m, src, _ in parts:
                if m == "gather" and src.requires_grad:
                    csr_cache.prefetch(idx, src.shape[2])
        return group_assemble(parts, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grads = []
        c0 = 0
        for n, mode in enumerate(ctx.modes):
            C, N = ctx.shapes[n]
            g = grad_out[:, c0:c0 + C]
            c0 += C
            gs = gc = None
            if mode == "gather":
                if ctx.needs_input_grad[2 + 2 * n]:
                    gcont = g.contiguous()
                    rec = _log.begin("group_bwd", grad_out=gcont, idx=idx, N=N) if _log.on else None
                    gs = group_bwd_auto(gcont, idx, N)
                    if rec is not None:
                        _log.end(rec, grad_f=gs)
                if ctx.has_center[n] and ctx.needs_input_grad[3 + 2 * n]:
                    gc = -g.sum(-1)
            elif ctx.needs_input_grad[2 + 2 * n]:
                gs = g.sum(-1)
            grads += [gs, gc]
        return (None, None, *grads)


class EdgeAffine(torch.autograd.Function):
    """p[idx] + LeakyReLU(q[idx] - center) (K12): the k-expanded half of the restructured EdgeConv
    (gcn_lib/pointnet/gcn.py:206-211).  Differentiable w.r.t. p, q, center."""

    @staticmethod
    def forward(ctx, p, q, center, idx, slope):
        ctx.slope = float(slope)
        ctx.save_for_backward(q, center, idx)
        if csr_cache.prefetch_enabled and (p.requires_grad or q.requires_grad):
            csr_cache.prefetch(idx, q.shape[2])
        return edge_affine_fwd(p, q, center, idx, slope)

    @staticmethod
    def backward(ctx, grad_out):
        q, center, idx = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        N = q.shape[2]
        need_p, need_q, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gp = gq = gc = None
        if need_p:
            gp = group_bwd_auto(grad_out, idx, N)
        if need_q or need_c:
            g2, gc = edge_affine_bwd(grad_out, q, center, idx, ctx.slope, want_center=need_c)
            if need_q:
                gq = group_bwd_auto(g2, idx, N)
        return gp, gq, gc, None, None


class ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.m = features.shape[2]
        ctx.save_for_backward(idx, weight)
        return three_interpolate_fwd(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        rec = _log.begin("three_interpolate_bwd", grad_out=grad_out, idx=idx, w=weight, m=ctx.m) if _log.on else None
        off, items = csr_cache.get(idx, ctx.m)
        gf = three_interpolate_bwd(grad_out, weight, off, items, ctx.m)
        if rec is not None:
            _log.end(rec, grad_f=gf)
        return gf, None, None


class ChamferSums(torch.autograd.Function):
    """Per-cloud Chamfer sums (sum_src [B], sum_tgt [B]); reductions over the batch stay
    in torch so every chamferdist reduction mode maps onto the same two kernels."""

    @staticmethod
    def forward(ctx, src, tgt, directions):
        rec = _log.begin("chamfer", src=src, tgt=tgt, directions=int(directions)) if _log.on else None
        r = chamfer_fwd(src, tgt, directions)
        if rec is not None:
            _log.end(rec, **r)
        ctx.directions = directions
        z = torch.zeros((src.shape[0],), dtype=torch.float32, device=src.device)
        i_s = r["i_src"] if r["i_src"] is not None else torch.empty(0, dtype=torch.int32, device=src.device)
        i_t = r["i_tgt"] if r["i_tgt"] is not None else torch.empty(0, dtype=torch.int32, device=src.device)
        ctx.save_for_backward(src, tgt, i_s, i_t)
        return (r["sum_src"] if r["sum_src"] is not None else z), (r["sum_tgt"] if r["sum_tgt"] is not None else z)

    @staticmethod
    def backward(ctx, g_src, g_tgt):
        src, tgt, i_s, i_t = ctx.saved_tensors
        need_src, need_tgt = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_src or need_tgt):
            return None, None, None
        g_src, g_tgt = g_src.contiguous(), g_tgt.contiguous()
        i_s, i_t = (i_s if i_s.numel() else None), (i_t if i_t.numel() else None)
        rec = _log.begin("chamfer_bwd", src=src, tgt=tgt, i_src=i_s, i_tgt=i_t, g_src=g_src, g_tgt=g_tgt,
                         directions=int(ctx.directions)) if _log.on else None
        gs, gt = chamfer_bwd(src, tgt, i_s, i_t, g_src, g_tgt, ctx.directions, need_src, need_tgt)
        if rec is not None:
            _log.end(rec, grad_src=gs, grad_tgt=gt)
        return gs, gt, None
