"""floydwarshall_b200 -- B200-native max-times Floyd-Warshall behind the API of
jinilover/floydWarshall's matrix-optimisation path (reference
src/lib/Algorithms.hs).  The compute lives in csrc/ (CUDA, sm_100a) behind the
C ABI of include/fwgpu.h; this package is the host-side mirror of the
reference interface plus ctypes plumbing."""
from . import _lib  # noqa: F401
from ._lib import Context, FwError  # noqa: F401
