"""nfllib_b200 — B200-native NTT / pointwise hot path of NFLlib.

The product is the C-ABI shared library `libnflgpu.so` (sources in nfllib_b200/csrc, interface in
include/nflgpu.h) plus the C++11 drop-in header include/nfl_b200.hpp.  This Python package is plumbing for
tests and bench.py only: a ctypes binding of the C ABI (nfllib_b200.capi).  There is no CPU fallback: importing
`capi` without a built library raises, and every compute call without a CUDA device returns an error."""
from .capi import Context, NflGpuError, lib, lib_path, params, params_limits, DTYPES  # noqa: F401
