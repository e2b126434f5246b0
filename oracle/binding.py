"""ctypes bindings for the CPU oracle (liboracle.so) and the per-width real reference
(oracle/_ref/libref_wNN.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (shared_simd_scan_b200/) never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)


def _ptr(a: np.ndarray, typ=C.c_void_p):
    return a.ctypes.data_as(typ)


def build_oracle() -> str:
    """Compile liboracle.so if it is missing or stale (gcc, a second or two)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    hdr = os.path.join(HERE, "oracle.h")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    L = C.CDLL(build_oracle())
    sz = C.c_size_t
    L.orc_compressed_buffer_size.restype = sz
    L.orc_compressed_buffer_size.argtypes = [C.c_int, sz]
    L.orc_decompression_output_buffer_size.restype = sz
    L.orc_decompression_output_buffer_size.argtypes = [sz]
    L.orc_scan_output_buffer_size.restype = sz
    L.orc_scan_output_buffer_size.argtypes = [sz]
    L.orc_pack.restype = None
    L.orc_pack.argtypes = [C.c_void_p, sz, C.c_int, C.c_void_p]
    L.orc_unpack.restype = None
    L.orc_unpack.argtypes = [C.c_void_p, sz, C.c_int, C.c_void_p]
    L.orc_code_at.restype = C.c_uint32
    L.orc_code_at.argtypes = [C.c_void_p, sz, C.c_int]
    L.orc_scan_range.restype = sz
    L.orc_scan_range.argtypes = [C.c_void_p, sz, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
                                 C.POINTER(C.c_uint64)]
    L.orc_positions.restype = sz
    L.orc_positions.argtypes = [C.c_void_p, sz, C.c_uint64, C.c_void_p, sz]
    L.orc_shared_scan.restype = None
    L.orc_shared_scan.argtypes = [C.c_void_p, sz, C.c_int, C.c_void_p, C.c_void_p, sz, C.c_int, C.c_void_p, sz,
                                  C.c_void_p]
    L.orc_shared_scan_linear.restype = None
    L.orc_shared_scan_linear.argtypes = [C.c_void_p, sz, C.c_int, C.c_void_p, C.c_void_p, sz, C.c_int,
                                         C.c_void_p, C.c_void_p]
    L.orc_bitvec_combine.restype = C.c_uint64
    L.orc_bitvec_combine.argtypes = [C.c_int, C.c_void_p, C.c_void_p, sz, C.c_void_p]
    L.orc_unpack_at.restype = None
    L.orc_unpack_at.argtypes = [C.c_void_p, C.c_int, C.c_void_p, sz, C.c_void_p]
    L.orc_mix64.restype = C.c_uint64
    L.orc_mix64.argtypes = [C.c_uint64]
    L.orc_generate_codes.restype = None
    L.orc_generate_codes.argtypes = [C.c_void_p, sz, C.c_int, C.c_uint64, C.c_uint64]
    L.orc_time_scan_range.restype = C.c_double
    L.orc_time_scan_range.argtypes = [C.c_void_p, sz, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                      C.POINTER(C.c_uint64)]
    return L


# ---- numpy-level helpers --------------------------------------------------------------------

def compressed_buffer_size(w: int, n: int) -> int:
    return lib().orc_compressed_buffer_size(w, n)


def scan_output_buffer_size(n: int) -> int:
    return lib().orc_scan_output_buffer_size(n)


def decompression_output_buffer_size(n: int) -> int:
    return lib().orc_decompression_output_buffer_size(n)


def generate_codes(n: int, w: int, seed: int = 42, first_index: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().orc_generate_codes(_ptr(out), n, w, seed & (2**64 - 1), first_index)
    return out


def pack(codes, w: int) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    out = np.empty(compressed_buffer_size(w, len(codes)), dtype=np.uint8)
    lib().orc_pack(_ptr(codes), len(codes), w, _ptr(out))
    return out


def unpack(packed: np.ndarray, n: int, w: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int32)
    lib().orc_unpack(_ptr(packed), n, w, _ptr(out))
    return out


def scan_range(packed: np.ndarray, n: int, w: int, lo: int, hi: int, granule: int = 1):
    """-> (bitvector bytes actually written, count)"""
    out = np.zeros(scan_output_buffer_size(n), dtype=np.uint8)
    cnt = C.c_uint64(0)
    nb = lib().orc_scan_range(_ptr(packed), n, w, lo, hi, granule, _ptr(out), C.byref(cnt))
    return out[:nb].copy(), int(cnt.value)


def positions(bitvec: np.ndarray, n: int, base: int = 0) -> np.ndarray:
    cap = int(np.unpackbits(bitvec[: (n + 7) // 8]).sum()) + 1
    out = np.empty(cap, dtype=np.uint64)
    k = lib().orc_positions(_ptr(np.ascontiguousarray(bitvec)), n, base, _ptr(out), cap)
    return out[:k].copy()


def shared_scan(packed: np.ndarray, n: int, w: int, lo, hi, granule: int = 1):
    """-> (K x bytes_written bitvectors, counts[K])"""
    lo = np.ascontiguousarray(lo, dtype=np.uint32)
    hi = np.ascontiguousarray(hi, dtype=np.uint32)
    K = len(lo)
    stride = scan_output_buffer_size(n)
    outs = np.zeros((K, stride), dtype=np.uint8)
    counts = np.zeros(K, dtype=np.uint64)
    lib().orc_shared_scan(_ptr(packed), n, w, _ptr(lo), _ptr(hi), K, granule, _ptr(outs), stride, _ptr(counts))
    nb = 4 * ((n + 31) // 32) if granule == 32 else (n + 7) // 8
    return outs[:, :nb].copy(), counts


def shared_scan_linear(packed: np.ndarray, n: int, w: int, lo, hi, granule: int = 1):
    """-> (rows x K byte matrix, counts[K])"""
    lo = np.ascontiguousarray(lo, dtype=np.uint32)
    hi = np.ascontiguousarray(hi, dtype=np.uint32)
    K = len(lo)
    rows = 4 * ((n + 31) // 32) if granule == 32 else (n + 7) // 8
    out = np.zeros((rows, K), dtype=np.uint8)
    counts = np.zeros(K, dtype=np.uint64)
    lib().orc_shared_scan_linear(_ptr(packed), n, w, _ptr(lo), _ptr(hi), K, granule, _ptr(out), _ptr(counts))
    return out, counts


COMBINE = {"and": 0, "or": 1, "andnot": 2, "xor": 3}


def bitvec_combine(op: str, a: np.ndarray, b: np.ndarray, n: int):
    out = np.zeros(4 * ((n + 31) // 32), dtype=np.uint8)
    hits = lib().orc_bitvec_combine(COMBINE[op], _ptr(np.ascontiguousarray(a)), _ptr(np.ascontiguousarray(b)), n,
                                    _ptr(out))
    return out, int(hits)


def unpack_at(packed: np.ndarray, w: int, positions) -> np.ndarray:
    positions = np.ascontiguousarray(positions, dtype=np.uint64)
    out = np.empty(len(positions), dtype=np.int32)
    lib().orc_unpack_at(_ptr(packed), w, _ptr(positions), len(positions), _ptr(out))
    return out


def time_scan_range(packed: np.ndarray, n: int, w: int, lo: int, hi: int, threads: int, reps: int):
    cnt = C.c_uint64(0)
    s = lib().orc_time_scan_range(_ptr(packed), n, w, lo, hi, threads, reps, C.byref(cnt))
    return s, int(cnt.value)


# ---- the real reference, one .so per width ------------------------------------------------------

# widths at which each reference family is exact when BITS_NEEDED is substituted
SSE_EXACT = set(range(1, 27)) | {28, 32}
AVX_EXACT = set(range(1, 17))
SCALAR_EXACT = set(range(1, 17))

DECOMPRESS = {"unvectorized": 0, "128_sweep": 1, "128_nosweep": 2, "128_9bit": 3, "128": 4, "128_unrolled": 5,
              "128_aligned": 6, "256": 7, "256_avx2": 8}
SCAN = {"unvectorized": 0, "128": 1, "128_unrolled": 2, "256": 3, "256_unrolled": 4}
SHARED = {"128_sequential": 0, "128_sequential_unrolled": 1, "128_threaded": 2, "128_standard": 3,
          "128_standard_unrolled": 4, "128_parallel": 5, "256_sequential": 6, "256_standard": 7, "256_parallel": 8}
LINEAR = {"linear_standard": 0, "linear_simple": 1}


def ref_path(w: int) -> str:
    return os.path.join(HERE, "_ref", f"libref_w{w:02d}.so")


def ref_available(w: int) -> bool:
    return os.path.exists(ref_path(w))


@lru_cache(maxsize=None)
def ref(w: int) -> C.CDLL:
    # RTLD_LOCAL: every width exports the same symbol names
    L = C.CDLL(ref_path(w), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    sz = C.c_size_t
    L.ref_bits.restype = C.c_int
    L.ref_compressed_buffer_size.restype = sz
    L.ref_compressed_buffer_size.argtypes = [sz]
    L.ref_scan_output_buffer_size.restype = sz
    L.ref_scan_output_buffer_size.argtypes = [sz]
    L.ref_decompression_output_buffer_size.restype = sz
    L.ref_decompression_output_buffer_size.argtypes = [sz]
    L.ref_compress.restype = sz
    L.ref_compress.argtypes = [C.c_void_p, sz, C.c_void_p, sz]
    L.ref_decompress.restype = C.c_int
    L.ref_decompress.argtypes = [C.c_int, C.c_void_p, sz, C.c_void_p]
    L.ref_scan.restype = C.c_long
    L.ref_scan.argtypes = [C.c_int, C.c_int, C.c_void_p, sz, C.c_void_p]
    L.ref_shared_scan.restype = C.c_int
    L.ref_shared_scan.argtypes = [C.c_int, C.c_void_p, sz, C.c_void_p, sz, C.c_void_p]
    L.ref_shared_scan_linear.restype = C.c_int
    L.ref_shared_scan_linear.argtypes = [C.c_int, C.c_void_p, sz, C.c_void_p, sz, C.c_void_p]
    L.ref_time_scan.restype = C.c_double
    L.ref_time_scan.argtypes = [C.c_int, C.c_int, C.c_void_p, sz, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    L.ref_time_decompress.restype = C.c_double
    L.ref_time_decompress.argtypes = [C.c_int, C.c_void_p, sz, C.c_int, C.c_int]
    L.ref_time_shared_scan.restype = C.c_double
    L.ref_time_shared_scan.argtypes = [C.c_int, C.c_void_p, sz, C.c_void_p, sz, C.c_int, C.c_int]
    L.ref_time_shared_scan_linear.restype = C.c_double
    L.ref_time_shared_scan_linear.argtypes = [C.c_int, C.c_void_p, sz, C.c_void_p, sz, C.c_int, C.c_int]
    assert L.ref_bits() == w
    return L


def ref_compress(codes, w: int) -> np.ndarray:
    """compress_9bit_input of the width-w reference build (uint16 input, w <= 16)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint16)
    n = len(codes)
    out = np.zeros(ref(w).ref_compressed_buffer_size(n), dtype=np.uint8)
    ref(w).ref_compress(_ptr(codes), n, _ptr(out), out.size)
    return out


def ref_decompress(variant: str, packed: np.ndarray, n: int, w: int) -> np.ndarray:
    # +128 ints beyond the reference's own 32-byte padding: decompress_unvectorized keeps writing up
    # to the end of the last 64-bit word when code n-1 straddles two words (the `oi == input_size`
    # exit at src/simd_scan_decompression.cpp:30 is skipped by the straddle branch :42-48), which
    # overruns decompression_output_buffer_size(n) for small widths (seen at w=3, n=1003).
    out = np.zeros(ref(w).ref_decompression_output_buffer_size(n) // 4 + 128, dtype=np.int32)
    rc = ref(w).ref_decompress(DECOMPRESS[variant], _ptr(packed), n, _ptr(out))
    assert rc == 0, variant
    return out[:n].copy()


def ref_scan(variant: str, key: int, packed: np.ndarray, n: int, w: int):
    """-> (whole scan_output_buffer_size(n) output vector, hits)"""
    out = np.zeros(ref(w).ref_scan_output_buffer_size(n), dtype=np.uint8)
    hits = ref(w).ref_scan(SCAN[variant], key, _ptr(packed), n, _ptr(out))
    assert hits >= 0, variant
    return out, int(hits)


def ref_shared_scan(variant: str, keys, packed: np.ndarray, n: int, w: int) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    sz = ref(w).ref_scan_output_buffer_size(n)
    outs = np.zeros((len(keys), sz), dtype=np.uint8)
    rc = ref(w).ref_shared_scan(SHARED[variant], _ptr(keys), len(keys), _ptr(packed), n, _ptr(outs))
    assert rc == 0, variant
    return outs


def ref_shared_scan_linear(variant: str, keys, packed: np.ndarray, n: int, w: int) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    out = np.zeros(len(keys) * ref(w).ref_scan_output_buffer_size(n), dtype=np.uint8)
    rc = ref(w).ref_shared_scan_linear(LINEAR[variant], _ptr(keys), len(keys), _ptr(packed), n, _ptr(out))
    assert rc == 0, variant
    return out
