"""Operator-level Python mirror of the reference's entry points (src/simd_scan.hpp:46-120).

torch is used for what it is good at here -- device memory, streams, pinned host buffers -- and
nothing else: every operator is one call into the C ABI (include/sss_b200.h), which launches the
hand-written sm_100a kernels.  Names follow the reference:

    reference (C++)                               here
    compress_9bit_input(vector<uint16_t>)         compress(codes, w)
    decompress_*(input, n, int* out)              decompress(packed, n, w)
    scan_*(key, input, n, vector<uint8_t>& out)   scan(packed, n, w, lo, hi)         (lo == hi == key)
    shared_scan_*(keys, input, n, outputs)        shared_scan(packed, n, w, lo, hi)
    shared_scan_128_linear_*(keys, input, n, out) shared_scan_linear(packed, n, w, lo, hi)
    (none; north star)                            positions(bitvec, n)

All device operators are asynchronous on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

_contexts = {}


class Context:
    """One sss_ctx per device (staging slots for the host entry points, workspaces)."""

    def __init__(self, device: int):
        self.device = device
        h = C.c_void_p()
        check(lib().sss_ctx_create(device, C.byref(h)), "sss_ctx_create")
        self.handle = h

    @staticmethod
    def current() -> "Context":
        if not torch.cuda.is_available():
            raise RuntimeError("shared_simd_scan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.cuda.current_device()
        ctx = _contexts.get(dev)
        if ctx is None:
            ctx = _contexts[dev] = Context(dev)
        return ctx

    @property
    def sm_count(self) -> int:
        return lib().sss_ctx_sm_count(self.handle)

    def close(self):
        if self.handle:
            lib().sss_ctx_destroy(self.handle)
            self.handle = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def packed_bytes(n: int, w: int) -> int:
    """compressed_buffer_size (src/simd_scan.hpp:20-26)."""
    return lib().sss_packed_bytes(n, w)


def bitvec_bytes(n: int) -> int:
    """scan_output_buffer_size (src/simd_scan.hpp:35-40)."""
    return lib().sss_bitvec_bytes(n)


def _check_packed(packed: torch.Tensor, n: int, w: int):
    if not (packed.is_cuda and packed.dtype == torch.uint8 and packed.is_contiguous()):
        raise ValueError("packed must be a contiguous CUDA uint8 tensor")
    if packed.numel() < packed_bytes(n, w):
        raise ValueError(f"packed holds {packed.numel()} bytes, needs packed_bytes(n, w) = {packed_bytes(n, w)}")


def _bounds(lo, hi, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """K query bounds as uint32 bit patterns in device memory.

    CUDA tensors with 4-byte elements are used as they are (no copy, no sync); anything else goes
    through a host uint32 array (K <= 1024, so this is tiny) and one H2D copy."""
    def conv(x):
        if isinstance(x, torch.Tensor) and x.is_cuda and x.element_size() == 4 and x.is_contiguous():
            return x
        if isinstance(x, torch.Tensor):
            x = x.cpu().numpy()
        arr = np.ascontiguousarray(np.asarray(x, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        return torch.from_numpy(arr.view(np.int32)).to(device)
    lo_t, hi_t = conv(lo), conv(hi)
    if lo_t.numel() != hi_t.numel() or lo_t.numel() < 1 or lo_t.numel() > _lib.SSS_MAX_PREDICATES:
        raise ValueError("lo and hi must have the same length K, 1 <= K <= 1024")
    return lo_t, hi_t, lo_t.numel()


# ---- device operators ---------------------------------------------------------------------------

def compress(codes: torch.Tensor, w: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """compress_9bit_input for any width: CUDA int32/uint32 codes -> packed uint8 buffer."""
    ctx = Context.current()
    if not (codes.is_cuda and codes.is_contiguous() and codes.element_size() == 4):
        raise ValueError("codes must be a contiguous CUDA tensor of 32-bit integers")
    n = codes.numel()
    if out is None:
        out = torch.empty(packed_bytes(n, w), dtype=torch.uint8, device=codes.device)
    check(lib().sss_pack(ctx.handle, codes.data_ptr(), n, w, out.data_ptr(), _stream()), "sss_pack")
    return out


def generate_packed(n: int, w: int, seed: int = 42, first_index: int = 0,
                    out: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """Synthetic uniform column, generated directly in packed form on the device."""
    ctx = Context.current()
    if out is None:
        out = torch.empty(packed_bytes(n, w), dtype=torch.uint8, device=device or "cuda")
    check(lib().sss_generate_packed(ctx.handle, out.data_ptr(), n, w, seed & (2**64 - 1), first_index, _stream()),
          "sss_generate_packed")
    return out


def decompress(packed: torch.Tensor, n: int, w: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """decompress_*: one int32 per code."""
    ctx = Context.current()
    _check_packed(packed, n, w)
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=packed.device)
    check(lib().sss_unpack(ctx.handle, packed.data_ptr(), n, w, out.data_ptr(), _stream()), "sss_unpack")
    return out


def scan(packed: torch.Tensor, n: int, w: int, lo: int, hi: Optional[int] = None,
         out: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None
         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """scan_*: range predicate lo <= code <= hi (hi defaults to lo: the reference's equality scan).

    Returns (bitvector uint8[bitvec_bytes(n)], count int64[1]); bits >= n are zero."""
    ctx = Context.current()
    _check_packed(packed, n, w)
    hi = lo if hi is None else hi
    if out is None:
        out = torch.empty(bitvec_bytes(n), dtype=torch.uint8, device=packed.device)
    if count is None:
        count = torch.empty(1, dtype=torch.int64, device=packed.device)
    check(lib().sss_scan_range(ctx.handle, packed.data_ptr(), n, w, lo & 0xFFFFFFFF, hi & 0xFFFFFFFF,
                               out.data_ptr(), count.data_ptr(), _stream()), "sss_scan_range")
    return out, count


def positions(bitvec: torch.Tensor, n: int, base: int = 0, capacity: Optional[int] = None,
              out: Optional[torch.Tensor] = None, num: Optional[torch.Tensor] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ascending position list of a bitvector -> (positions int64[capacity], num int64[1])."""
    ctx = Context.current()
    if capacity is None:
        capacity = n if out is None else out.numel()
    pos = out if out is not None else torch.empty(max(capacity, 1), dtype=torch.int64, device=bitvec.device)
    if pos.numel() < capacity:
        raise ValueError("out holds fewer than `capacity` positions")
    if num is None:
        num = torch.empty(1, dtype=torch.int64, device=bitvec.device)
    check(lib().sss_positions(ctx.handle, bitvec.data_ptr(), n, base, pos.data_ptr(), capacity, num.data_ptr(),
                              _stream()), "sss_positions")
    return pos, num


def scan_positions(packed: torch.Tensor, n: int, w: int, lo: int, hi: Optional[int] = None, base: int = 0,
                   capacity: Optional[int] = None, out: Optional[torch.Tensor] = None,
                   pos_out: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The north star's scan output in one call: bitvector + match count + ascending position list
    (each position plus `base`) -> (bitvector, positions int64[capacity], count int64[1])."""
    ctx = Context.current()
    _check_packed(packed, n, w)
    hi = lo if hi is None else hi
    if out is None:
        out = torch.empty(bitvec_bytes(n), dtype=torch.uint8, device=packed.device)
    if capacity is None:
        capacity = n
    pos = pos_out if pos_out is not None else torch.empty(max(capacity, 1), dtype=torch.int64, device=packed.device)
    if pos.numel() < capacity:
        raise ValueError("pos_out holds fewer than `capacity` positions")
    if count is None:
        count = torch.empty(1, dtype=torch.int64, device=packed.device)
    check(lib().sss_scan_range_positions(ctx.handle, packed.data_ptr(), n, w, lo & 0xFFFFFFFF, hi & 0xFFFFFFFF,
                                         out.data_ptr(), count.data_ptr(), base, pos.data_ptr(), capacity, _stream()),
          "sss_scan_range_positions")
    return out, pos, count


def shared_scan(packed: torch.Tensor, n: int, w: int, lo, hi=None, out: Optional[torch.Tensor] = None,
                counts: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """shared_scan_*: K predicates in one pass -> (uint8[K, stride] bitvectors, int64[K] counts)."""
    ctx = Context.current()
    _check_packed(packed, n, w)
    lo32, hi32, K = _bounds(lo, lo if hi is None else hi, packed.device)
    if out is None:
        stride = (bitvec_bytes(n) + 127) // 128 * 128  # whole 128-byte lines per warp store (+3 % at K = 64)
        out = torch.empty((K, stride), dtype=torch.uint8, device=packed.device)
    if out.dim() != 2 or out.shape[0] != K or not out.is_contiguous():
        raise ValueError("out must be a contiguous [K, stride] uint8 tensor")
    if counts is None:
        counts = torch.empty(K, dtype=torch.int64, device=packed.device)
    check(lib().sss_shared_scan(ctx.handle, packed.data_ptr(), n, w, lo32.data_ptr(), hi32.data_ptr(), K,
                                out.data_ptr(), out.stride(0), counts.data_ptr(), _stream()), "sss_shared_scan")
    return out, counts


def shared_scan_linear(packed: torch.Tensor, n: int, w: int, lo, hi=None, out: Optional[torch.Tensor] = None,
                       counts: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """shared_scan_128_linear_*: -> (uint8[rows, K] with rows = 4*ceil(n/32), int64[K] counts)."""
    ctx = Context.current()
    _check_packed(packed, n, w)
    lo32, hi32, K = _bounds(lo, lo if hi is None else hi, packed.device)
    rows = 4 * ((n + 31) // 32)
    if out is None:
        out = torch.empty((rows, K), dtype=torch.uint8, device=packed.device)
    if out.numel() < rows * K or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 tensor of at least 4*ceil(n/32)*K bytes")
    if counts is None:
        counts = torch.empty(K, dtype=torch.int64, device=packed.device)
    check(lib().sss_shared_scan_linear(ctx.handle, packed.data_ptr(), n, w, lo32.data_ptr(), hi32.data_ptr(), K,
                                       out.data_ptr(), counts.data_ptr(), _stream()), "sss_shared_scan_linear")
    return out, counts


_OPS = {"and": 0, "or": 1, "andnot": 2, "xor": 3}


def bitvec_combine(op: str, a: torch.Tensor, b: torch.Tensor, n: int, out: Optional[torch.Tensor] = None,
                   count: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """a OP b over the first n bits (op in and / or / andnot / xor) -> (bitvector, popcount int64[1])."""
    ctx = Context.current()
    if out is None:
        out = torch.empty(bitvec_bytes(n), dtype=torch.uint8, device=a.device)
    if count is None:
        count = torch.empty(1, dtype=torch.int64, device=a.device)
    check(lib().sss_bitvec_combine(ctx.handle, _OPS[op], a.data_ptr(), b.data_ptr(), n, out.data_ptr(),
                                   count.data_ptr(), _stream()), "sss_bitvec_combine")
    return out, count


def decompress_at(packed: torch.Tensor, w: int, positions: torch.Tensor, num: Optional[int] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Late materialisation: the codes of a packed column at a position list (int64 / uint64 CUDA tensor)."""
    ctx = Context.current()
    num = positions.numel() if num is None else num
    if out is None:
        out = torch.empty(max(num, 1), dtype=torch.int32, device=packed.device)
    check(lib().sss_unpack_at(ctx.handle, packed.data_ptr(), w, positions.data_ptr(), num, out.data_ptr(), _stream()),
          "sss_unpack_at")
    return out[:num]


_CALIB = {"read": 0, "write": 1, "copy": 2, "mix": 3, "streams32": 4, "streams128": 5}


def calibrate(mode: str, a: Optional[torch.Tensor], b: Optional[torch.Tensor], nbytes: int, ratio: int = 1) -> None:
    """One streaming pass with the access mix `mode` (read / write / copy / mix = read nbytes, write
    nbytes/ratio); the caller times it.  See sss_calibrate."""
    ctx = Context.current()
    check(lib().sss_calibrate(ctx.handle, _CALIB[mode], a.data_ptr() if a is not None else None,
                              b.data_ptr() if b is not None else None, nbytes, ratio, _stream()), "sss_calibrate")


# ---- host-buffer operators (numpy arrays or pinned CPU tensors) ------------------------------------

Host = Union[np.ndarray, torch.Tensor]


def _hptr(a: Host) -> int:
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.is_contiguous()
        return a.data_ptr()
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def host_scan(packed: Host, n: int, w: int, lo: int, hi: Optional[int] = None, tail_granule: int = 1,
              out: Optional[Host] = None) -> Tuple[Host, int]:
    """Host-buffer scan (H2D + kernel + D2H inside the call); see sss_host_scan_range."""
    ctx = Context.current()
    hi = lo if hi is None else hi
    if out is None:
        out = np.zeros(bitvec_bytes(n), dtype=np.uint8)
    cnt = C.c_uint64(0)
    check(lib().sss_host_scan_range(ctx.handle, _hptr(packed), n, w, lo & 0xFFFFFFFF, hi & 0xFFFFFFFF, tail_granule,
                                    _hptr(out), C.addressof(cnt)), "sss_host_scan_range")
    return out, int(cnt.value)


def host_decompress(packed: Host, n: int, w: int, out: Optional[Host] = None) -> Host:
    ctx = Context.current()
    if out is None:
        out = np.empty(n, dtype=np.int32)
    check(lib().sss_host_unpack(ctx.handle, _hptr(packed), n, w, _hptr(out)), "sss_host_unpack")
    return out


def host_compress(codes: np.ndarray, w: int) -> np.ndarray:
    ctx = Context.current()
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    out = np.empty(packed_bytes(len(codes), w), dtype=np.uint8)
    check(lib().sss_host_pack(ctx.handle, _hptr(codes), len(codes), w, _hptr(out)), "sss_host_pack")
    return out


def host_shared_scan(packed: Host, n: int, w: int, lo: Sequence[int], hi: Optional[Sequence[int]] = None,
                     tail_granule: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    ctx = Context.current()
    lo_a = np.ascontiguousarray(np.asarray(lo, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
    hi_a = lo_a if hi is None else np.ascontiguousarray(np.asarray(hi, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
    K = len(lo_a)
    stride = bitvec_bytes(n)
    out = np.zeros((K, stride), dtype=np.uint8)
    counts = np.zeros(K, dtype=np.uint64)
    check(lib().sss_host_shared_scan(ctx.handle, _hptr(packed), n, w, _hptr(lo_a), _hptr(hi_a), K, tail_granule,
                                     _hptr(out), stride, _hptr(counts)), "sss_host_shared_scan")
    return out, counts


def host_shared_scan_linear(packed: Host, n: int, w: int, lo: Sequence[int], hi: Optional[Sequence[int]] = None,
                            tail_granule: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    ctx = Context.current()
    lo_a = np.ascontiguousarray(np.asarray(lo, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
    hi_a = lo_a if hi is None else np.ascontiguousarray(np.asarray(hi, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
    K = len(lo_a)
    out = np.zeros(K * bitvec_bytes(n), dtype=np.uint8)  # test/simd_scan_tests.cpp:119 sizing
    counts = np.zeros(K, dtype=np.uint64)
    check(lib().sss_host_shared_scan_linear(ctx.handle, _hptr(packed), n, w, _hptr(lo_a), _hptr(hi_a), K,
                                            tail_granule, _hptr(out), _hptr(counts)), "sss_host_shared_scan_linear")
    return out, counts


def host_shared_scan_into(packed: Host, n: int, w: int, lo: np.ndarray, hi: np.ndarray, out: Host, stride: int,
                          counts: np.ndarray, tail_granule: int = 1) -> None:
    """sss_host_shared_scan into caller-owned (for full PCIe speed: pinned) buffers: `out` holds K
    bitvectors `stride` bytes apart, `counts` K uint64."""
    ctx = Context.current()
    K = len(lo)
    check(lib().sss_host_shared_scan(ctx.handle, _hptr(packed), n, w, _hptr(lo), _hptr(hi), K, tail_granule,
                                     _hptr(out), stride, _hptr(counts)), "sss_host_shared_scan")


def pinned_empty(nbytes: int) -> torch.Tensor:
    """Pinned host buffer (full-speed PCIe copies for the host_* operators)."""
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def bind_host_thread(device: int) -> int:
    """Bind the calling thread to the CPUs next to GPU `device` (sss_bind_host_thread): pinned buffers
    allocated afterwards are node-local.  Returns the number of CPUs the thread may run on."""
    n = C.c_int(0)
    check(lib().sss_bind_host_thread(device, C.byref(n)), "sss_bind_host_thread")
    return n.value
