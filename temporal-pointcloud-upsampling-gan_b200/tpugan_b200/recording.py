"""Boundary-call log of the kernel front end (``tpugan_b200.functional``).

Two uses, both off by default and free when off:

* ``capture``: every boundary call keeps device-side clones of its inputs and outputs, so a
  run of the reference's unmodified train step on these kernels can afterwards be checked
  *call by call* against the CPU oracle on each call's own recorded inputs (the dense layers
  between the calls then cannot blur the comparison).  Clones are taken on the launching
  stream, nothing synchronises.
* ``timing``: a CUDA event pair around every call -> device time of the hot path inside a
  full train step (the rest is cuDNN / elementwise / optimiser work of the model).

The log is process-global like the kernels' launch counter; it is test / bench
instrumentation of the product path, not a second path.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch


class Call:
    __slots__ = ("op", "inputs", "outputs", "start", "stop")

    def __init__(self, op):
        self.op = op
        self.inputs: Dict[str, Any] = {}
        self.outputs: Dict[str, Any] = {}
        self.start = None
        self.stop = None

    def device_ms(self) -> float:
        return self.start.elapsed_time(self.stop)


class CallLog:
    def __init__(self):
        self.on = False
        self.capture = False
        self.timing = False
        self.calls: List[Call] = []

    def start(self, capture: bool = True, timing: bool = False) -> None:
        self.calls = []
        self.capture, self.timing = capture, timing
        self.on = True

    def stop(self) -> List[Call]:
        self.on = False
        calls, self.calls = self.calls, []
        return calls

    # -- used by functional.py -------------------------------------------------------------
    def begin(self, _name: str, **inputs) -> Optional[Call]:
        """Called before the kernels of a boundary call are launched."""
        if not self.on:
            return None
        c = Call(_name)
        if self.capture:
            c.inputs = {k: _keep(v) for k, v in inputs.items()}
        if self.timing:
            c.start = torch.cuda.Event(enable_timing=True)
            c.start.record()
        return c

    def end(self, c: Optional[Call], **outputs) -> None:
        if c is None:
            return
        if self.timing:
            c.stop = torch.cuda.Event(enable_timing=True)
            c.stop.record()
        if self.capture:
            c.outputs = {k: _keep(v) for k, v in outputs.items()}
        self.calls.append(c)


def _keep(v):
    if isinstance(v, torch.Tensor):
        return v.detach().clone()
    if isinstance(v, (list, tuple)):
        return [_keep(x) for x in v]
    return v


log = CallLog()


def hot_path_ms(calls) -> Dict[str, float]:
    """Summed device time per op of a timed log (call after a synchronize)."""
    out: Dict[str, float] = {}
    for c in calls:
        if c.start is not None:
            out[c.op] = out.get(c.op, 0.0) + c.device_ms()
    return out
