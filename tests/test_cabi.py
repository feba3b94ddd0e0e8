"""CPU: the C-ABI library loads, exports every symbol include/ihpr_b200.h declares, and its argument
validation works without a GPU (no compute is launched here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ihpr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ihpr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(built_lib):
    from ihpr_b200 import _lib
    handle = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(handle, name), "library does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_version_and_variant_knob(built_lib):
    import ihpr_b200
    assert ihpr_b200.version() == 100
    ihpr_b200.set_variant(2)
    assert ihpr_b200.get_variant() == 2
    ihpr_b200.set_variant(0)


def test_workspace_bytes(built_lib):
    from ihpr_b200._lib import lib
    L = lib()
    a = L.ihpr_workspace_bytes(1, 18, 64, 64, 64)
    b = L.ihpr_workspace_bytes(32, 18, 64, 64, 64)
    assert 0 < a < b < 64 << 20 and a % 256 == 0
    assert L.ihpr_workspace_bytes(0, 18, 64, 64, 64) == 0


def test_argument_validation_without_gpu(built_lib):
    from ihpr_b200._lib import lib
    L = lib()
    rc = L.ihpr_softargmax3d_fwd(None, 0, 1, 2, 4, 4, 4, None, None, None, 0, None)
    assert rc == -1 and b"null" in L.ihpr_last_error()
    rc = L.ihpr_softargmax3d_fwd(None, 7, 1, 2, 4, 4, 4, None, None, None, 0, None)
    assert rc == -1 and b"dtype" in L.ihpr_last_error()
    rc = L.ihpr_integral_l1_bwd(None, 0, 0, 2, 4, 4, 4, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"non-positive" in L.ihpr_last_error()


def test_fused_head_argument_validation_without_gpu(built_lib):
    from ihpr_b200._lib import lib
    L = lib()
    one = 1     # any non-null pointer: validation happens before anything is dereferenced
    assert L.ihpr_head_softargmax_fwd(None, None, None, 1, 256, 18, 64, 64, 64, None, None, None) == -1
    assert L.ihpr_head_softargmax_fwd(16, 16, 16, 1, 200, 18, 64, 64, 64, 16, None, None) == -1 and b"multiple of 64" in L.ihpr_last_error()
    assert L.ihpr_head_softargmax_fwd(16, 16, 16, 1, 256, 18, 48, 64, 64, 16, None, None) == -1 and b"depth_dim" in L.ihpr_last_error()
    assert L.ihpr_head_softargmax_fwd(16, 16, 16, 1, 256, 18, 64, 60, 60, 16, None, None) == -1 and b"W %" in L.ihpr_last_error()
    assert L.ihpr_head_integral_l1_bwd(16, 16, 16, 1, 256, 18, 64, 64, 64, 16, None, None, None, None, None, None, None, None) == -1
    assert L.ihpr_integral_l1_fwd_bwd(None, 0, 1, 2, 4, 4, 4, None, None, None, None, None, None, None, None, 0, None) == -1
    assert L.ihpr_scale_grad(None, 0, 16, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    from ihpr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", "/nonexistent/libihpr_b200.so")
    with pytest.raises(_lib.IhprError, match="not built"):
        _lib.lib()
