"""libzseek_b200 — B200-native reader for libzseek seekable files (zstd / LZ4 frames + seek table).

The product is the C-ABI shared library ``libzseek_b200.so`` (C host reader + hand-written sm_100a
kernels, see ``csrc/``) that exports the reference's reader API (include/zseek.h) and the additive
GPU entry points (include/zseek_b200.h).  This package is the Python-side mirror of that interface
(ctypes; same names, argument meaning and error behaviour) used by the tests and the benchmark.
There is no CPU decode path: importing works anywhere, but opening a reader without a usable
sm_100 device raises.
"""
from .reader import (LIB_PATH, Reader, ReaderStats, ZseekError, build, load_library, pread_full)

__all__ = ["LIB_PATH", "Reader", "ReaderStats", "ZseekError", "build", "load_library", "pread_full"]
