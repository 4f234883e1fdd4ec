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


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The descriptor structs are passed by pointer across the boundary, so the ctypes mirrors in _C.py must have the C compiler's
    layout for include/hupr_b200.h: compile a C program that prints sizeof / offsetof of every mirrored field and compare."""
    import shutil
    import subprocess
    from hupr_b200 import _C
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    mirrors = {"hupr_conv_desc": _C.ConvDesc, "hupr_attn_desc": _C.AttnDesc, "hupr_attn_bwd_desc": _C.AttnBwdDesc,
               "hupr_wgrad_desc": _C.WgradDesc, "hupr_tensor_view": _C.TensorView, "hupr_pack_job": _C.PackJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hupr_b200.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for field, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, field, cname, field))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        if not line:
            continue
        cname, field, value = line.split()
        cls = mirrors[cname]
        if field == "sizeof":
            assert ctypes.sizeof(cls) == int(value), "%s: ctypes size %d, C size %s" % (cname, ctypes.sizeof(cls), value)
        else:
            assert getattr(cls, field).offset == int(value), "%s.%s" % (cname, field)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())
    # and the header has no field the mirror lacks: the C struct would then be larger than the last mirrored field's end
    for cname, cls in mirrors.items():
        last, ctype = cls._fields_[-1]
        assert getattr(cls, last).offset + ctypes.sizeof(ctype) + 8 > ctypes.sizeof(cls)
