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


class Recorder:
    """Collects (op, inputs, outputs) of every boundary call while enabled.  With
    shapes_only=True arrays are reduced to {shape, dtype} (call schedules for bench.py)."""

    def __init__(self):
        self.enabled = False
        self.shapes_only = False
        self.calls = []
        self.next_id = 0

    def new_id(self):
        self.next_id += 1
        return self.next_id

    def record(self, op, inputs, outputs):
        if not self.enabled:
            return
        if self.shapes_only:
            inputs = {k: _describe(v) for k, v in inputs.items()}
            outputs = {k: _describe(v) for k, v in outputs.items()}
        self.calls.append((op, inputs, outputs))

    def start(self, shapes_only=False):
        self.calls = []
        self.next_id = 0
        self.shapes_only = shapes_only
        self.enabled = True

    def stop(self):
        self.enabled = False
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
