"""kwage_b200: B200-native (sm_100a) implementation of KWAGE's k-mer Bloom-filter hot path.

The product is libkwage_cuda.so behind the C ABI in include/kwage_cuda.h plus the C++ host layer in
kwage_b200/host/ that mirrors the reference's make_bloom_filter / build_db / search.  This Python
package is only a ctypes binding used by the tests and the benchmark."""
from . import capi  # noqa: F401

__all__ = ["capi"]
