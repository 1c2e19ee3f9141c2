"""Oracle-backed CPU shims + call recorder (test infrastructure; see README.md)."""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def _describe(v):
    import numpy as np

    if isinstance(v, np.ndarray):
        return {"shape": list(v.shape), "dtype": str(v.dtype)}
    if isinstance(v, (np.floating, float)):
        return float(v)
    if isinstance(v, (np.integer, int)):
        return int(v)
    return v


class DepTracker:
    """Data-flow tracker for schedule recording (test infrastructure).  While active, every aten op that
    runs on the CPU reference propagates, per tensor storage, the set of boundary calls whose results the
    data was computed from (only the most recent boundary call along each path: a boundary call's outputs
    restart the set).  make_schedule.py stores the sets with every call so that a replay may overlap
    independent chains (e.g. the three frames of a window) without inventing concurrency."""

    def __init__(self):
        import torch
        from torch.utils._python_dispatch import TorchDispatchMode
        from torch.utils._pytree import tree_leaves

        self.torch = torch
        self.deps = {}  # storage address -> frozenset(call indices); overwritten when an address is re-allocated
        tracker = self

        class Mode(TorchDispatchMode):
            def __torch_dispatch__(self, func, types, args=(), kwargs=None):
                kwargs = kwargs or {}
                out = func(*args, **kwargs)
                tin = [a for a in tree_leaves((args, kwargs)) if isinstance(a, torch.Tensor)]
                in_keys = {tracker.key(t) for t in tin}
                ins = frozenset().union(*[tracker.deps.get(k, frozenset()) for k in in_keys]) if in_keys else frozenset()
                for o in tree_leaves(out):
                    if isinstance(o, torch.Tensor):
                        k = tracker.key(o)
                        if k in in_keys:
                            tracker.deps[k] = tracker.deps.get(k, frozenset()) | ins  # view / in-place result
                        else:
                            tracker.deps[k] = ins                                       # fresh allocation
                schema = getattr(func, "_schema", None)
                if schema is not None and ins:
                    for a, sa in zip(args, schema.arguments):
                        if sa.alias_info is not None and sa.alias_info.is_write:
                            for t in tree_leaves(a):
                                if isinstance(t, torch.Tensor):
                                    k = tracker.key(t)
                                    tracker.deps[k] = tracker.deps.get(k, frozenset()) | ins
                return out

        self.mode = Mode()

    @staticmethod
    def key(t):
        try:
            return t.untyped_storage().data_ptr()
        except Exception:
            return id(t)

    def of(self, t):
        return self.deps.get(self.key(t), frozenset())

    def tag(self, t, call_index):
        self.deps[self.key(t)] = frozenset([call_index])


class Recorder:
    """Collects (op, inputs, outputs) of every boundary call while enabled.  With
    shapes_only=True arrays are reduced to {shape, dtype} (call schedules for bench.py)."""

    def __init__(self):
        self.enabled = False
        self.shapes_only = False
        self.calls = []
        self.next_id = 0
        self.tracker = None

    def new_id(self):
        self.next_id += 1
        return self.next_id

    def record(self, op, inputs, outputs):
        if not self.enabled:
            return
        inputs = {k: v for k, v in inputs.items() if not (k in ("deps", "p1_sha", "p2_sha") and v is None)}
        if self.shapes_only:
            inputs = {k: _describe(v) for k, v in inputs.items()}
            outputs = {k: _describe(v) for k, v in outputs.items()}
        self.calls.append((op, inputs, outputs))

    def deps(self, **tensors):
        """{input name: sorted boundary-call indices its data derives from} (None when not tracking)"""
        if self.tracker is None or not self.enabled:
            return None
        return {k: sorted(self.tracker.of(t)) for k, t in tensors.items() if t is not None and hasattr(t, "untyped_storage")}

    def tag(self, *tensors):
        """the given tensors are outputs of the call recorded last"""
        if self.tracker is None or not self.enabled:
            return
        for t in tensors:
            self.tracker.tag(t, len(self.calls) - 1)

    def start(self, shapes_only=False, track_deps=False):
        self.calls = []
        self.next_id = 0
        self.shapes_only = shapes_only
        self.enabled = True
        self.tracker = None
        if track_deps:
            self.tracker = DepTracker()
            self.tracker.mode.__enter__()

    def stop(self):
        self.enabled = False
        if self.tracker is not None:
            self.tracker.mode.__exit__(None, None, None)
            self.tracker = None
        return self.calls


recorder = Recorder()


def activate():
    """Put the oracle-backed packages first on sys.path (CPU reference runs only)."""
    for name in ("pytorch3d", "pytorch3d.ops", "frnn", "pointnet2_ops", "pointnet2_ops.pointnet2_utils",
                 "chamferdist", "dgl", "dgl.utils", "dgl.function", "dgl.nn", "dgl.geometry"):
        sys.modules.pop(name, None)
    if SHIM_DIR in sys.path:
        sys.path.remove(SHIM_DIR)
    sys.path.insert(0, SHIM_DIR)
    return SHIM_DIR
