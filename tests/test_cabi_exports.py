"""The C-ABI library loads and exports every symbol include/damvs.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "damvs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(damvs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from damvsnet_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("damvs_warp_agg_fwd", "damvs_conv3d_fwd", "damvs_softmax_regress_fwd", "damvs_homo_warp_fwd",
                 "damvs_depth_regression_fwd", "damvs_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/damvs.h but not exported"


def test_binding_table_matches_header():
    from damvsnet_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_abi_version_and_error_string(lib):
    lib.damvs_abi_version.restype = ctypes.c_int
    assert lib.damvs_abi_version() == 1
    lib.damvs_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.damvs_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    """Bad arguments are rejected on the host before any CUDA call."""
    lib.damvs_softmax_regress_fwd.restype = ctypes.c_int
    rc = lib.damvs_softmax_regress_fwd(None, None, None, None, None, None, 1, 8, 4, 4, 1, None)
    assert rc == 1
    lib.damvs_last_error.restype = ctypes.c_char_p
    assert b"null pointer" in lib.damvs_last_error()


def test_sass_is_sm100a_only(lib):
    import subprocess
    from damvsnet_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
