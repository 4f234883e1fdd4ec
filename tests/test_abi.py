"""CPU tests of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol that
include/hupr_b200.h declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hupr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hupr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    handle = ctypes.CDLL(built_lib)
    names = declared_symbols()
    assert "hupr_fft_cascade_i16" in names and "hupr_version" in names
    for name in names:
        assert hasattr(handle, name), "libhupr_b200.so does not export %s" % name


def test_binder_covers_header(built_lib):
    from hupr_b200 import _C
    assert sorted(_C.SIGNATURES) == declared_symbols()
    lib = _C.lib()
    assert lib.hupr_version() >= 100
    assert lib.hupr_error_string(0) == b"ok"
    assert b"sm_100" in lib.hupr_error_string(-4)


def test_no_cpu_fallback_on_missing_library(monkeypatch, built_lib):
    from hupr_b200 import _C
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/libhupr_b200.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _C.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hupr-a-benchmark-for-human-pose-estimation-using-millimeter-wave-radar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
