"""opflow_b200 -- B200-native evaluation engine for OpFlow's stencil hot path.

The product is `libopflow_b200.so` (CUDA kernels for sm_100a behind the C ABI in include/opflow_b200.h) plus the C++
front-end headers in opflow_b200/include (`#include <OpFlow>`).  This package is the ctypes view of the same ABI used by
tests/ and bench.py.  Importing it does not load the library; the first call does, and fails loudly when the library is
not built or no CUDA device is visible (there is no CPU path).
"""
from . import capi, host  # noqa: F401

__all__ = ["capi", "host"]
