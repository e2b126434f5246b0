"""shared_simd_scan_b200 -- B200-native (sm_100a) fused bit-unpack + (shared) range scan.

A from-scratch CUDA implementation of the one hot path of RRr89/Shared_SIMD_Scan, behind the
reference's own entry points:

  include/sss_b200.h        the C ABI (libsss_b200.so) -- the drop-in boundary
  include/simd_scan.hpp     the reference's C++ signatures re-created on top of it
  shared_simd_scan_b200.api the same operators for Python callers (torch tensors as device memory)
  shared_simd_scan_b200.dist range-partitioned multi-GPU scan (one process per GPU, NCCL count allreduce)

Importing the package does not need a GPU; calling any operator does (no CPU fallback).
"""
from . import _lib  # noqa: F401
from .api import (  # noqa: F401
    Context,
    bitvec_bytes,
    compress,
    decompress,
    generate_packed,
    packed_bytes,
    positions,
    scan,
    shared_scan,
    shared_scan_linear,
)

__all__ = ["Context", "bitvec_bytes", "compress", "decompress", "generate_packed", "packed_bytes", "positions",
           "scan", "shared_scan", "shared_scan_linear"]
