"""The C++ host-side mirror (floydwarshall_b200/host/algorithms.hpp) replays the reference's
AlgorithmsTest.hs through the C ABI: host-only cases on CPU, everything on the GPU."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "test_algorithms")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")], stdout=subprocess.DEVNULL)


def test_cpp_host_cpu_cases():
    _build()
    out = subprocess.run([EXE, "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_gpu_cases():
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


MULTI = os.path.join(HERE, "cpp", "test_multi")


def test_cpp_multi_cpu_cases():
    """fw_multi_plan (the schedule as data) and argument checks from C++, no device."""
    _build()
    out = subprocess.run([MULTI, "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [2, 4])
def test_cpp_multi_gpu_cases(ndev):
    """fw_multi_* from C++ on ndev shards (virtual ranks when the box has fewer GPUs): the reference's golden
    answers through fw_multi_sync + fw_multi_optimum, and fw_multi_solve_edges == fw_solve_edges."""
    _build()
    out = subprocess.run([MULTI, "gpu", str(ndev)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
