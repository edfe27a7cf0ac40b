"""minotert_b200 -- B200-native MinoteRT ray-tracing hot path.

The product is libminotert.so (CUDA, sm_100a; sources in csrc/, C ABI in include/minotert.h).
This package holds only the ctypes binding, the host-side mirror of the reference's
src/gfx interface, and the synthetic scene generators used by tests and bench.py.
"""
from . import capi, scenes  # noqa: F401
from .capi import Context, MinoteError  # noqa: F401
