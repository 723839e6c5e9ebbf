"""ellalgo_rs_b200 — B200-native ellipsoid update path behind the ellalgo-rs trait API.

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of include/ellcuda.h, built to
csrc/libellcuda.so).  This package is the thin host mirror used by tests and bench.  There is no CPU
fallback: compute calls raise when the CUDA library or a B200 is missing.
"""
from . import _abi  # noqa: F401
from .problem import BatchOutput, BatchProblem  # noqa: F401
