"""Drop-in for ``chamferdist`` (krrish94/chamferdist >= 1.0) as imported by the reference:
``from chamferdist import ChamferDistance`` — loss.py:3, used at loss.py:125,176,224,280."""
from .chamfer import ChamferDistance  # noqa: F401

__all__ = ["ChamferDistance"]
