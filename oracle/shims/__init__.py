"""Oracle-backed CPU shims + call recorder (test infrastructure; see README.md)."""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


class Recorder:
    """Collects (op, inputs, outputs) of every boundary call while enabled."""

    def __init__(self):
        self.enabled = False
        self.calls = []

    def record(self, op, inputs, outputs):
        if self.enabled:
            self.calls.append((op, inputs, outputs))

    def start(self):
        self.calls = []
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
