"""ctypes binding of ``libtpugan_b200.so`` (the C ABI declared in ``include/tpugan_b200.h``).

There is deliberately **no fallback**: if the shared library is missing or a call
returns a non-zero status this module raises.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p, c_long

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libtpugan_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

TPG_OK = 0
ABI_VERSION = 9
TPG_EINVAL, TPG_EUNSUPPORTED, TPG_ECUDA, TPG_EWORKSPACE = -1, -2, -3, -4
REDUCE_MAX, REDUCE_SUM, REDUCE_MIN = 0, 1, 2
CHAMFER_FWD, CHAMFER_REV, CHAMFER_BOTH = 1, 2, 3
PART_GATHER, PART_BROADCAST = 0, 1
ASSEMBLE_MAX_PARTS = 4


class AssemblePart(ctypes.Structure):
    """struct tpg_assemble_part (include/tpugan_b200.h)."""
    _fields_ = [("src", c_void_p), ("center", c_void_p), ("C", c_int), ("N", c_int), ("mode", c_int)]


class TpgError(RuntimeError):
    """A libtpugan_b200 call failed (message from tpg_last_error())."""


class TpgLibraryMissing(ImportError):
    pass


_P, _I, _F, _Z = c_void_p, c_int, c_float, c_size_t

# name -> (restype, argtypes); mirrors include/tpugan_b200.h one to one
_PROTOS = {
    "tpg_abi_version": (_I, []),
    "tpg_last_error": (c_char_p, []),
    "tpg_launch_count": (c_uint64, []),
    "tpg_set_option": (_I, [c_char_p, c_long]),
    "tpg_knn_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "tpg_knn_fallback_count_offset": (_Z, [_I]),
    "tpg_knn_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "tpg_knn_cond_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P, _P]),
    "tpg_bytes_equal_and": (_I, [_P, _P, _Z, _P, _P]),
    "tpg_knn_take_prefix": (_I, [_P, _P, _P, _I, _P, _P, _I, ctypes.c_longlong, _P]),
    "tpg_frnn_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "tpg_frnn_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _Z, _P]),
    "tpg_ball_query_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "tpg_ball_query_f32": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P, _Z, _P]),
    "tpg_fps_workspace_bytes": (_Z, [_I, _I]),
    "tpg_fps_f32": (_I, [_P, _I, _I, _I, _P, _P, _Z, _P]),
    "tpg_fps_start_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "tpg_group_fwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "tpg_group_fwd_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "tpg_group_fwd_ws_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "tpg_inverse_index_workspace_bytes": (_Z, [_I, _I, _I]),
    "tpg_inverse_index_build": (_I, [_P, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "tpg_group_bwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "tpg_group_bwd_segments": (_I, [_I, _I, _I, _I]),
    "tpg_group_bwd_segmented_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "tpg_group_reduce_fwd_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "tpg_group_reduce_workspace_bytes": (_Z, [_I, _I, _I]),
    "tpg_group_reduce_fwd_ws_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "tpg_group_reduce_bwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "tpg_group_assemble_f32": (_I, [_P, _I, _P, _I, _I, _I, _P, _P]),
    "tpg_edge_affine_fwd_f32": (_I, [_P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P, _P]),
    "tpg_edge_affine_bwd_f32": (_I, [_P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P, _P, _P]),
    "tpg_three_nn_workspace_bytes": (_Z, [_I, _I, _I]),
    "tpg_three_nn_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "tpg_three_interpolate_fwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "tpg_three_interpolate_bwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "tpg_chamfer_fwd_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "tpg_chamfer_fwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "tpg_chamfer_bwd_workspace_bytes": (_Z, [_I, _I, _I]),
    "tpg_chamfer_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "tpg_cubic_interp_workspace_bytes": (_Z, [_I, _I, _I]),
    "tpg_cubic_interp_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _Z, _P]),
    "tpg_gather_rows_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "tpg_gather_rows_bwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "tpg_knn_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises TpgLibraryMissing when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TpgLibraryMissing(
            f"{LIB_NAME} not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C temporal-pointcloud-upsampling-gan_b200/csrc` "
            f"(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError == ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.tpg_abi_version() != ABI_VERSION:
        raise TpgLibraryMissing(f"{LIB_NAME}: ABI version {lib.tpg_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().tpg_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def launch_count() -> int:
    return int(load().tpg_launch_count())


def check(status: int, what: str) -> None:
    if status != TPG_OK:
        raise TpgError(f"{what} failed (status {status}): {last_error()}")


_fns = {}


def call(name: str, *args) -> None:
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(load(), name)
    status = fn(*args)
    if status != TPG_OK:
        check(status, name)


def set_option(name: str, value: int) -> None:
    """Scheduling hint of the library (include/tpugan_b200.h: tpg_set_option); never changes a result."""
    call("tpg_set_option", name.encode(), int(value))
