"""arcflow_b200 — B200-native (sm_100a) implementation of the ArcFlow denoising hot path.

Package layout (only what the path needs):
  csrc/      hand-written CUDA (tcgen05 GEMM + attention, streaming kernels, engine) + the C ABI
  _lib.py    ctypes mirror of include/arcflow_b200.h
  ops.py     tensor-level operators (one kernel each)
  build.py   in-tree nvcc build of lib/libarcflow_b200.so
"""
from ._lib import AfbError, LIB_PATH  # noqa: F401

__all__ = ["AfbError", "LIB_PATH"]
