"""B200-native (sm_100a) kernel layer for the fp64 shallow-water step of ocean_model_arch.

The product is libswcuda.so (CUDA kernels + C ABI, include/swcuda.h); this package is the thin
host-side mirror of the reference's algorithm layer used by the tests and the benchmark.
"""
from . import _lib  # noqa: F401
from ._lib import MODE_FUSED, MODE_REFERENCE, SwcuError  # noqa: F401
