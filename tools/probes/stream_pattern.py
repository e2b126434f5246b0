#!/usr/bin/env python
"""Developer probe: HBM write bandwidth of K interleaved output streams (the bitvector layout's store
pattern, no computation) with 32-bit and 128-bit stores, next to a single write stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from shared_simd_scan_b200 import api  # noqa: E402

nbytes = 8 << 30
b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: api.calibrate("write", None, b, nbytes))
print(f"single stream, 128-bit stores: {nbytes / ms / 1e6:8.1f} GB/s")
a = torch.empty(nbytes // 4 + 4096, dtype=torch.uint8, device="cuda")
a.random_(0, 256)
for K in (1, 8, 64, 256, 1024):
    for mode in ("streams32", "streams128"):
        ms = timed(lambda: api.calibrate(mode, None, b, nbytes, K))
        # with reads: 64 bytes read per thread and tile (K x 4 bytes written): the traffic of K predicates over 16-bit codes
        rd = nbytes // 4 // K * 64          # 64 bytes read per stream word (per thread and tile)
        if rd > a.numel():
            print(f"K = {K:5d} {mode:11s}: write only {nbytes / ms / 1e6:8.1f} GB/s")
            continue
        ms_r = timed(lambda: api.calibrate(mode, a, b, nbytes, K))
        print(f"K = {K:5d} {mode:11s}: write only {nbytes / ms / 1e6:8.1f} GB/s   with reads {(nbytes + rd) / ms_r / 1e6:8.1f} GB/s "
              f"(reads {100 * rd / (nbytes + rd):.0f} %)")
