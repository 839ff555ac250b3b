"""The C-ABI library loads on a CPU-only box and exports every function include/fwgpu.h declares
(no compute calls here)."""
import ctypes
import os
import re

from floydwarshall_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fwgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(fw_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fwgpu.h but not exported by libfwgpu.so"
    assert sorted(_lib.SYMBOLS) == names, "floydwarshall_b200/_lib.py binds a different set than the header declares"


def test_library_reports_version_and_has_no_fallback():
    L = _lib.load()
    assert b"sm_100a" in L.fw_version()
    if L.fw_device_count() == 0:
        h = ctypes.c_void_p()
        rc = L.fw_ctx_create(0, ctypes.byref(h))
        assert rc == _lib.FW_ERR_CUDA and b"no CPU fallback" in L.fw_last_error()
