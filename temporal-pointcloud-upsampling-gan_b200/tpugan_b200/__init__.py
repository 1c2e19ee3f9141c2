"""tpugan_b200 — B200-native (sm_100a) point-neighbourhood kernels behind TPU-GAN's call surfaces.

The directory that contains this package is a *site directory*: next to
``tpugan_b200`` it holds drop-in packages named exactly like the native
dependencies the reference imports (``pytorch3d.ops``, ``frnn``,
``pointnet2_ops.pointnet2_utils``, ``chamferdist``).  Put it on ``sys.path``
(``tpugan_b200.activate()`` does that) and the reference's unmodified model, loss
and train-step code runs on these kernels.

No CPU fallback exists: importing is cheap, but the first op call loads
``libtpugan_b200.so`` and raises if it was not built.
"""
from __future__ import annotations

import os
import sys

from . import _lib
from ._lib import TpgError, TpgLibraryMissing, launch_count, set_option  # noqa: F401

__version__ = "0.1.0"

SITE_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB_DIR = os.path.join(SITE_DIR, "import_stubs")


def activate(import_stubs: bool = False) -> str:
    """Make the drop-in packages importable (they take precedence over any installed
    package of the same name).  ``import_stubs=True`` additionally exposes import-only
    stand-ins for packages the reference imports but never calls on the train step
    (``dgl``, ``emd``, ``open3d``, ``tensorboardX``)."""
    for d in ([STUB_DIR] if import_stubs else []) + [SITE_DIR]:
        if d in sys.path:
            sys.path.remove(d)
        sys.path.insert(0, d)
    return SITE_DIR


def patch_reference(mods=None, **kw):
    """Rebind the reference's thin wrappers (ball_query_wrapper, IDGCN gather+max, cubic_interpolation,
    interpolate_vel_lst) to the fused / batched kernels; see tpugan_b200.reference_patches."""
    from .reference_patches import patch_reference as _patch

    return _patch(mods, **kw)


def library_path() -> str:
    return _lib.LIB_PATH


def is_built() -> bool:
    return os.path.exists(_lib.LIB_PATH)
