"""Range-partitioned multi-GPU scan: one process per GPU (torch.distributed), no data-path collective.

The reference has no multi-device code; its only parallel variant partitions the PREDICATES over
OpenMP threads and re-reads the whole column once per predicate (src/simd_scan_shared.cpp:25-32).
Codes are independent, so here the COLUMN is range-partitioned instead:

  * rank r owns codes [first_r, first_r + count_r), with first_r a multiple of 128 codes, so every
    shard starts 16-byte aligned in the packed stream for any width and on a 16-byte boundary of
    every bitvector;
  * each rank scans its shard with the single-GPU kernels;
  * per-query match counts are combined with ONE small all-reduce (8 bytes per query; NCCL over
    NVLink on GPUs, gloo in the CPU tests);
  * position lists are rebased locally (global = first_r + local); concatenating the ranks' lists in
    rank order is globally ascending.  Nothing else crosses GPUs.

The per-shard operators default to the CUDA path (shared_simd_scan_b200.api).  They are parameters
only so that the host-side logic (partitioning, all-reduce, rebasing) can be exercised by the
world_size-2 gloo tests on a machine without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

SHARD_ALIGN = 128  # codes; 128 codes = w 16-byte vectors


def shard_range(n: int, world: int, rank: int, align: int = SHARD_ALIGN) -> Tuple[int, int]:
    """Contiguous range partition of n codes: (first, count) of `rank`.  Every shard but the last is
    a multiple of `align` codes; trailing ranks may be empty when n is small."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-n // world)               # ceil
    per = -(-per // align) * align     # round up to the alignment
    first = min(rank * per, n)
    return first, min(per, n - first)


@dataclass
class Shard:
    """This rank's slice of a globally range-partitioned packed column."""
    packed: object      # device buffer holding codes [first, first+count) in packed form
    n_global: int
    first: int
    count: int
    w: int


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def generate_shard(n_global: int, w: int, seed: int = 42) -> Shard:
    """Synthetic global column: every rank generates exactly its own range of the counter stream."""
    from . import api
    world, rank = _world()
    first, count = shard_range(n_global, world, rank)
    return Shard(api.generate_packed(count, w, seed=seed, first_index=first), n_global, first, count, w)


def _allreduce_counts(counts: torch.Tensor) -> torch.Tensor:
    world, _ = _world()
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def scan(shard: Shard, lo: int, hi: Optional[int] = None, scan_op: Optional[Callable] = None):
    """-> (this rank's slice of the bitvector, GLOBAL match count as a 1-element int64 tensor)."""
    if scan_op is None:
        from . import api
        scan_op = api.scan
    bitvec, count = scan_op(shard.packed, shard.count, shard.w, lo, lo if hi is None else hi)
    return bitvec, _allreduce_counts(count)


def shared_scan(shard: Shard, lo, hi=None, linear: bool = False, shared_op: Optional[Callable] = None):
    """-> (this rank's slice of the K bitvectors / linear matrix, GLOBAL per-query counts int64[K])."""
    if shared_op is None:
        from . import api
        shared_op = api.shared_scan_linear if linear else api.shared_scan
    out, counts = shared_op(shard.packed, shard.count, shard.w, lo, lo if hi is None else hi)
    return out, _allreduce_counts(counts)


def positions(shard: Shard, bitvec, capacity: Optional[int] = None, positions_op: Optional[Callable] = None):
    """-> (this rank's positions rebased to global code indices, number of local matches)."""
    if positions_op is None:
        from . import api
        positions_op = api.positions
    return positions_op(bitvec, shard.count, base=shard.first, capacity=capacity)


def gather_positions(local_positions: torch.Tensor, local_num: int) -> torch.Tensor:
    """Convenience for small results: concatenate every rank's rebased list in rank order (globally
    ascending).  This is a result-collection helper, not part of the scan path."""
    world, _ = _world()
    mine = local_positions[:local_num].contiguous()
    if world == 1:
        return mine
    sizes = [torch.zeros(1, dtype=torch.int64, device=mine.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_num], dtype=torch.int64, device=mine.device))
    cap = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(cap, dtype=mine.dtype, device=mine.device)
    padded[:local_num] = mine
    parts = [torch.zeros(cap, dtype=mine.dtype, device=mine.device) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[: int(s.item())] for p, s in zip(parts, sizes)])
