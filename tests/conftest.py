import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")
    config.addinivalue_line("markers", "slow: CPU oracle work of a minute or more")


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a CPU-only box skips the gpu-marked tests instead of failing at the first one
    (tests/test_algorithms_api.py::test_no_cpu_fallback stays unmarked and checks the loud failure)."""
    try:
        from floydwarshall_b200 import _lib
        have_gpu = _lib.load().fw_device_count() > 0
    except Exception:  # noqa: BLE001  (library not built: the gpu tests could not run anyway)
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device (libfwgpu has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
