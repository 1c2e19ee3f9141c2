"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/tpugan_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tpugan_b200.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from tpugan_b200 import _lib

    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tpg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_python_binds(built):
    assert set(declared_functions()) == set(built.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name
    lib.tpg_abi_version.restype = ctypes.c_int
    assert lib.tpg_abi_version() == built.ABI_VERSION == 9  # host-only call


def test_library_is_sm100a_only(built):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", built.LIB_PATH], capture_output=True,
                         text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_host_only_entry_points_do_not_need_a_gpu(built):
    lib = built.load()
    assert lib.tpg_fps_workspace_bytes(4, 65536) == 0
    assert lib.tpg_fps_workspace_bytes(4, 65537) == 4 * 65537 * 4
    assert lib.tpg_inverse_index_workspace_bytes(2, 100, 1000) > 0
    assert lib.tpg_chamfer_bwd_workspace_bytes(2, 100, 200) > 0
    assert lib.tpg_cubic_interp_workspace_bytes(1, 100, 100) > 0
    assert lib.tpg_launch_count() >= 0
    # scheduling hints: host-only, validated
    assert lib.tpg_set_option(b"fps.sms_per_cloud", 1) == built.TPG_OK
    assert lib.tpg_set_option(b"fps.sms_per_cloud", 8) == built.TPG_OK
    assert lib.tpg_set_option(b"fps.sms_per_cloud", 3) == built.TPG_EINVAL
    assert lib.tpg_set_option(b"fps.exclusive_sm", 1) == built.TPG_OK
    assert lib.tpg_set_option(b"fps.exclusive_sm", 0) == built.TPG_OK
    assert lib.tpg_set_option(b"fps.exclusive_sm", 2) == built.TPG_EINVAL
    assert lib.tpg_set_option(b"no.such.option", 1) == built.TPG_EINVAL
    assert b"no.such.option" in lib.tpg_last_error()


def test_argument_errors_are_reported_not_thrown(built):
    lib = built.load()
    # bad arguments are rejected before any CUDA call
    assert lib.tpg_knn_f32(None, None, None, None, 1, 4, 4, 300, 1, None, None, None, 0, None) == built.TPG_EUNSUPPORTED
    assert b"D=300" in lib.tpg_last_error()
    assert lib.tpg_frnn_f32(None, None, None, None, 1, 4, 4, 5, 1, 0.1, None, None, None, None, 0, None) \
        == built.TPG_EUNSUPPORTED
    assert lib.tpg_group_reduce_fwd_f32(None, None, 1, 1, 1, 1, 1, 7, None, None, None) == built.TPG_EINVAL
    assert lib.tpg_knn_f32(None, None, None, None, 1, 4, 4, 3, 1, None, None, None, 0, None) == built.TPG_EINVAL  # null ptr


def test_product_fails_loudly_without_cuda(built):
    import torch

    from tpugan_b200 import functional as F

    if torch.cuda.is_available():
        pytest.skip("GPU box")
    with pytest.raises(RuntimeError, match="no CPU path"):
        F.knn(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(import oracle|from oracle)", txt, flags=re.M), os.path.join(d, f)
                assert "libtpg_oracle" not in txt, os.path.join(d, f)
