"""Single-process multi-GPU scans: a thin Python mirror of the sss_mgpu_* entry points of the C ABI
(include/sss_b200.h; csrc/mgpu.cu).  One process drives N GPUs, the column is range-partitioned at
multiples of 128 codes, the per-query counts are combined with one ncclAllReduce, position lists are
rebased by the shard's first code.  (shared_simd_scan_b200.dist is the one-process-per-GPU flavour of
the same partitioning on top of torch.distributed.)  torch is only the owner of the device buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def shard_range(n_global: int, n_shards: int, shard: int) -> Tuple[int, int]:
    first, count = C.c_uint64(), C.c_uint64()
    check(lib().sss_shard_range(n_global, n_shards, shard, C.byref(first), C.byref(count)), "sss_shard_range")
    return first.value, count.value


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


class MultiGpu:
    def __init__(self, devices: Optional[Sequence[int]] = None, n_devices: Optional[int] = None):
        if devices is None:
            n = n_devices if n_devices is not None else torch.cuda.device_count()
            devices = list(range(n))
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        check(lib().sss_mgpu_create(len(self.devices), arr, C.byref(h)), "sss_mgpu_create")
        self.handle = h

    def __len__(self):
        return len(self.devices)

    def close(self):
        if self.handle:
            lib().sss_mgpu_destroy(self.handle)
            self.handle = None

    def sync(self):
        check(lib().sss_mgpu_sync(self.handle), "sss_mgpu_sync")

    def shard(self, n_global: int, g: int) -> Tuple[int, int]:
        return shard_range(n_global, len(self.devices), g)

    def generate(self, n_global: int, w: int, seed: int = 42) -> List[torch.Tensor]:
        """The global synthetic column, every GPU generating exactly its own range of it."""
        out = []
        for g, dev in enumerate(self.devices):
            first, count = self.shard(n_global, g)
            t = torch.empty(lib().sss_packed_bytes(count, w), dtype=torch.uint8, device=f"cuda:{dev}")
            check(lib().sss_generate_packed(lib().sss_mgpu_ctx(self.handle, g), t.data_ptr(), count, w,
                                            seed & (2**64 - 1), first, lib().sss_mgpu_stream(self.handle, g)),
                  "sss_generate_packed")
            out.append(t)
        self.sync()
        return out

    def scan(self, packed: Sequence[torch.Tensor], n_global: int, w: int, lo: int, hi: Optional[int] = None,
             with_positions: bool = False):
        """-> (bitvector slices, global count per GPU [int64[1] each], positions per GPU or None, local counts)"""
        hi = lo if hi is None else hi
        G = len(self.devices)
        bv, cnt, pos, num, caps = [], [], [], [], (C.c_uint64 * G)()
        for g, dev in enumerate(self.devices):
            _, count = self.shard(n_global, g)
            d = f"cuda:{dev}"
            bv.append(torch.empty(lib().sss_bitvec_bytes(count), dtype=torch.uint8, device=d))
            cnt.append(torch.empty(1, dtype=torch.int64, device=d))
            num.append(torch.empty(1, dtype=torch.int64, device=d))
            pos.append(torch.empty(max(count, 1), dtype=torch.int64, device=d) if with_positions else None)
            caps[g] = count
        torch.cuda.synchronize()
        check(lib().sss_mgpu_scan_range(self.handle, _ptr_array(packed), n_global, w, lo & 0xFFFFFFFF, hi & 0xFFFFFFFF,
                                        _ptr_array(bv), _ptr_array(cnt), _ptr_array(pos) if with_positions else None,
                                        caps if with_positions else None, _ptr_array(num)), "sss_mgpu_scan_range")
        self.sync()
        return bv, cnt, (pos if with_positions else None), num

    def shared_scan(self, packed: Sequence[torch.Tensor], n_global: int, w: int, lo, hi, linear: bool = False):
        """-> (per-GPU outputs, per-GPU GLOBAL counts int64[K])"""
        lo_a = np.ascontiguousarray(np.asarray(lo, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        hi_a = np.ascontiguousarray(np.asarray(hi, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        K = len(lo_a)
        outs, cnts, stride = [], [], 0
        for g, dev in enumerate(self.devices):
            _, count = self.shard(n_global, g)
            d = f"cuda:{dev}"
            if linear:
                outs.append(torch.empty((4 * ((count + 31) // 32), K), dtype=torch.uint8, device=d))
            else:
                stride = (lib().sss_bitvec_bytes(self.shard(n_global, 0)[1]) + 127) // 128 * 128
                outs.append(torch.empty((K, stride), dtype=torch.uint8, device=d))
            cnts.append(torch.empty(K, dtype=torch.int64, device=d))
        torch.cuda.synchronize()
        check(lib().sss_mgpu_shared_scan(self.handle, _ptr_array(packed), n_global, w, lo_a.ctypes.data, hi_a.ctypes.data,
                                         K, 1 if linear else 0, _ptr_array(outs), stride, _ptr_array(cnts)),
              "sss_mgpu_shared_scan")
        self.sync()
        return outs, cnts

    def host_scan(self, packed: np.ndarray, n: int, w: int, lo: int, hi: Optional[int] = None, tail_granule: int = 1,
                  out=None):
        hi = lo if hi is None else hi
        if out is None:
            out = np.zeros(lib().sss_bitvec_bytes(n), dtype=np.uint8)
        cnt = C.c_uint64(0)
        ptr = lambda a: a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data  # noqa: E731
        check(lib().sss_mgpu_host_scan_range(self.handle, ptr(packed), n, w, lo & 0xFFFFFFFF, hi & 0xFFFFFFFF,
                                             tail_granule, ptr(out), C.addressof(cnt)), "sss_mgpu_host_scan_range")
        return out, int(cnt.value)

    def host_shared_scan(self, packed: np.ndarray, n: int, w: int, lo, hi, tail_granule: int = 1):
        lo_a = np.ascontiguousarray(np.asarray(lo, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        hi_a = np.ascontiguousarray(np.asarray(hi, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        K = len(lo_a)
        stride = lib().sss_bitvec_bytes(n)
        out = np.zeros((K, stride), dtype=np.uint8)
        counts = np.zeros(K, dtype=np.uint64)
        check(lib().sss_mgpu_host_shared_scan(self.handle, packed.ctypes.data, n, w, lo_a.ctypes.data, hi_a.ctypes.data, K,
                                              tail_granule, out.ctypes.data, stride, counts.ctypes.data),
              "sss_mgpu_host_shared_scan")
        return out, counts
