"""The C-ABI library loads on a CPU-only box and exports every symbol include/parafem_b200.h
declares; device entry points fail with a status code (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from parafem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "parafem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 45
    L = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    _lib.lib()


def test_xx3_compat_symbols_are_exported():
    """include/parafem_xx3_compat.h: the six names xx3.f90:56-148 binds today, so xx3 links unchanged."""
    txt = open(os.path.join(ROOT, "include", "parafem_xx3_compat.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = sorted(set(re.findall(r"\bint\s+([a-z_0-9]+)\s*\(", txt)))
    assert syms == sorted(_lib.XX3_SIGNATURES) and len(syms) == 6
    L = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), s


def test_signatures_are_plain_c():
    """No C++/torch types in the boundary: only C scalars, pointers and one POD struct."""
    txt = open(os.path.join(ROOT, "include", "parafem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)       # declarations only, comments stripped
    for bad in ("std::", "torch", "at::", "template", "class "):
        assert bad not in txt


def test_device_path_fails_loudly_without_gpu():
    import shutil
    if shutil.which("nvidia-smi") and os.system("nvidia-smi -L > /dev/null 2>&1") == 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    rc = _lib.lib().pf_init(0, 1, 0, None, C.byref(h))
    assert rc > 0
    buf = C.create_string_buffer(512)
    _lib.lib().pf_last_error(None, buf, 512)
    assert b"no CPU path" in buf.value


def test_product_never_touches_the_oracle():
    """parafem_b200/ must not import, link or dlopen anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "parafem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(l for l in txt.splitlines() if not l.strip().startswith(("#", "//", "*", '"""')))
                assert "import oracle" not in code and "from oracle" not in code and "libpf_oracle" not in code, f
    out = os.popen(f"ldd {_lib.LIB_PATH}").read()
    assert "oracle" not in out
