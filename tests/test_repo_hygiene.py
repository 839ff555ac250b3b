"""Source-tree rules that are cheap to check on every CPU run:
  * the product package never imports, links or executes anything under oracle/ (the oracle is the checker);
  * no tracked source names the batched-memcpy runtime API family (the GPU pool's gate refuses runs whose sources
    do; for many small copies use one async copy per piece, a gather kernel, or fewer and larger copies);
  * the multi-GPU entry fails loudly without a CUDA device, like every other compute entry (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from floydwarshall_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER_FILES = {"VERDICT.md", "ADVICE.md"}          # written by the round driver, not by this repository


def tracked_files():
    out = subprocess.run(["git", "ls-files"], cwd=ROOT, capture_output=True, text=True)
    if out.returncode != 0 or not out.stdout.strip():       # a snapshot without .git (GPU box): walk the tree instead
        files = []
        for d, dirs, fs in os.walk(ROOT):
            dirs[:] = [x for x in dirs if x not in (".git", "gpurun_out", "__pycache__", ".pytest_cache", ".hypothesis")]
            files += [os.path.relpath(os.path.join(d, f), ROOT) for f in fs]
        return files
    return out.stdout.split()


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "floydwarshall_b200")
    pat = re.compile(r"\boracle\b|fw_oracle|libfworacle")
    for d, _dirs, fs in os.walk(pkg):
        for f in fs:
            if not f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                continue
            text = open(os.path.join(d, f), errors="replace").read()
            code = [ln for ln in text.split("\n") if re.search(r"^\s*(import|from)\b|#include|dlopen|CDLL|subprocess", ln)]
            assert not any(pat.search(ln) for ln in code), f"{f} reaches into oracle/"


def test_no_batched_memcpy_api_names():
    name = re.compile("cuda" + "Memcpy" + "Batch|cu" + "Memcpy" + "Batch|Memcpy" + "3DBatch", re.I)
    for rel in tracked_files():
        if os.path.basename(rel) in DRIVER_FILES or rel.endswith((".so", ".ncu-rep", ".png")):
            continue
        path = os.path.join(ROOT, rel)
        if not os.path.isfile(path) or os.path.getsize(path) > 8 << 20:
            continue
        assert not name.search(open(path, errors="replace").read()), f"{rel} names a batched-memcpy API"


def test_multi_gpu_entry_has_no_cpu_fallback():
    L = _lib.load()
    if L.fw_device_count() > 0:
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = L.fw_multi_create(2, None, ctypes.byref(h))
    assert rc == _lib.FW_ERR_CUDA and not h.value
    rc = L.fw_multi_create_rank(0, 0, 1, None, ctypes.byref(h))
    assert rc == _lib.FW_ERR_CUDA and not h.value
