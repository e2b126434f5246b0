"""ctypes loader for libsss_b200.so (the C ABI declared in include/sss_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, the caller gets an exception -- never a silently slower path.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SSS_LIB_PATH: developer override to A/B-test an experimental build of the same ABI
LIB_PATH = os.environ.get("SSS_LIB_PATH") or os.path.join(PKG_DIR, "libsss_b200.so")

SSS_MAX_PREDICATES = 1024
TAIL_EXACT, TAIL_8, TAIL_32, TAIL_CURSOR32, TAIL_CURSOR64 = 1, 8, 32, 132, 164

# every symbol include/sss_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _i, _sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
SIGNATURES = {
    "sss_version": (_i, []),
    "sss_error_string": (C.c_char_p, [_i]),
    "sss_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "sss_ctx_destroy": (_i, [_vp]),
    "sss_ctx_sm_count": (_i, [_vp]),
    "sss_packed_bytes": (_sz, [_u64, _i]),
    "sss_unpack_bytes": (_sz, [_u64]),
    "sss_bitvec_bytes": (_sz, [_u64]),
    "sss_tail_eval_codes": (_u64, [_u64, _i, _i]),
    "sss_tail_out_bytes": (_u64, [_u64, _i, _i]),
    "sss_pack": (_i, [_vp, _vp, _u64, _i, _vp, _vp]),
    "sss_generate_packed": (_i, [_vp, _vp, _u64, _i, _u64, _u64, _vp]),
    "sss_unpack": (_i, [_vp, _vp, _u64, _i, _vp, _vp]),
    "sss_scan_range": (_i, [_vp, _vp, _u64, _i, _u32, _u32, _vp, _vp, _vp]),
    "sss_positions": (_i, [_vp, _vp, _u64, _u64, _vp, _u64, _vp, _vp]),
    "sss_scan_range_positions": (_i, [_vp, _vp, _u64, _i, _u32, _u32, _vp, _vp, _u64, _vp, _u64, _vp]),
    "sss_bitvec_combine": (_i, [_vp, _i, _vp, _vp, _u64, _vp, _vp, _vp]),
    "sss_unpack_at": (_i, [_vp, _vp, _i, _vp, _u64, _vp, _vp]),
    "sss_calibrate": (_i, [_vp, _i, _vp, _vp, _u64, _u32, _vp]),
    "sss_shared_scan": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _vp, _u64, _vp, _vp]),
    "sss_shared_scan_linear": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "sss_host_unpack": (_i, [_vp, _vp, _u64, _i, _vp]),
    "sss_host_scan_range": (_i, [_vp, _vp, _u64, _i, _u32, _u32, _i, _vp, _vp]),
    "sss_host_shared_scan": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _i, _vp, _u64, _vp]),
    "sss_host_shared_scan_ptrs": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "sss_host_shared_scan_linear": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "sss_host_pack": (_i, [_vp, _vp, _u64, _i, _vp]),
    "sss_host_positions": (_i, [_vp, _vp, _u64, _u64, _vp, _u64, _vp]),
    "sss_mgpu_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "sss_mgpu_destroy": (_i, [_vp]),
    "sss_mgpu_device_count": (_i, [_vp]),
    "sss_mgpu_device": (_i, [_vp, _i]),
    "sss_mgpu_ctx": (_vp, [_vp, _i]),
    "sss_mgpu_stream": (_vp, [_vp, _i]),
    "sss_mgpu_sync": (_i, [_vp]),
    "sss_shard_range": (_i, [_u64, _i, _i, C.POINTER(_u64), C.POINTER(_u64)]),
    "sss_mgpu_scan_range": (_i, [_vp, _vp, _u64, _i, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "sss_mgpu_shared_scan": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _i, _vp, _u64, _vp]),
    "sss_mgpu_host_scan_range": (_i, [_vp, _vp, _u64, _i, _u32, _u32, _i, _vp, _vp]),
    "sss_mgpu_host_shared_scan": (_i, [_vp, _vp, _u64, _i, _vp, _vp, _i, _i, _vp, _u64, _vp]),
    "sss_host_alloc": (_vp, [_sz]),
    "sss_host_free": (None, [_vp]),
    "sss_host_register": (_i, [_vp, _sz]),
    "sss_host_unregister": (_i, [_vp]),
    "sss_bind_host_thread": (_i, [_i, C.POINTER(_i)]),
}


class SSSError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what}: status {status} ({error_string(status)})")


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C shared_simd_scan_b200/csrc`). shared_simd_scan_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("SSS_LIB_PATH") and not hasattr(L, name):
            continue  # developer override: an older experimental build may lack the newest entry points
        fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


def error_string(status: int) -> str:
    try:
        return lib().sss_error_string(status).decode()
    except Exception:
        return "?"


def check(status: int, what: str) -> None:
    if status != 0:
        raise SSSError(status, what)
