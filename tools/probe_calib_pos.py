#!/usr/bin/env python
"""Developer probe: (1) sss_calibrate at 4 GiB and at the sizes of the 2^30-code scans at 1-3 bit (what can a
45-90 us kernel with that read:write mix reach?), (2) the fused scan + position list call against scan
followed by sss_positions.   python tools/probe_calib_pos.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def launch_ms(fn, reps=10, rounds=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


def main():
    big = 4 << 30
    a = torch.empty(big, dtype=torch.uint8, device="cuda")
    b = torch.empty(big, dtype=torch.uint8, device="cuda")
    a.random_(0, 256)
    for mode, ratio, nbytes in (("read", 1, big), ("write", 1, big), ("copy", 1, big), ("mix", 1, big), ("mix", 2, big),
                                ("mix", 3, big), ("mix", 4, big), ("mix", 8, big), ("mix", 17, big), ("mix", 32, big),
                                ("mix", 1, 128 << 20), ("mix", 2, 256 << 20), ("mix", 3, 384 << 20), ("mix", 4, 512 << 20),
                                ("copy", 1, 128 << 20)):
        ms = launch_ms(lambda: api.calibrate(mode, a, b, nbytes, ratio), reps=3 if nbytes == big else 10)
        moved = {"read": nbytes, "write": nbytes, "copy": 2 * nbytes}.get(mode, nbytes + nbytes // ratio)
        print(f"calib {mode:5s} ratio {ratio:2d} {nbytes >> 20:5d} MiB in: {ms * 1e3:8.1f} us {moved / ms / 1e6:7.1f} GB/s "
              f"{moved / ms / 1e6 / PEAK:.3f}")
    del a, b
    n = 1 << 30
    w = 17
    m = (1 << w) - 1
    col = api.generate_packed(n, w, seed=42 + w)
    bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    bv2 = torch.empty_like(bv)
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    num = torch.empty(1, dtype=torch.int64, device="cuda")
    for sel in (0.001, 0.01, 0.05, 0.25, 0.5):
        hi = max(0, int(sel * (m + 1)) - 1)
        api.scan(col, n, w, 0, hi, out=bv, count=cnt)
        hits = int(cnt.item())
        pos = torch.empty(max(hits, 1), dtype=torch.int64, device="cuda")
        pos2 = torch.empty_like(pos)
        api.positions(bv, n, capacity=hits, out=pos, num=num)
        ms_p = launch_ms(lambda: api.positions(bv, n, capacity=hits, out=pos, num=num), reps=5)
        ms_s = launch_ms(lambda: api.scan(col, n, w, 0, hi, out=bv2, count=cnt), reps=5)
        ms_f = launch_ms(lambda: api.scan_positions(col, n, w, 0, hi, capacity=hits, out=bv2, pos_out=pos2, count=cnt), reps=5)

        def two_calls():
            api.scan(col, n, w, 0, hi, out=bv, count=cnt)
            api.positions(bv, n, capacity=hits, out=pos, num=num)
        ms_2 = launch_ms(two_calls, reps=5)
        nb = 4 * ((n + 31) // 32)
        ok = (torch.equal(pos, pos2), torch.equal(bv[:nb], bv2[:nb]), int(cnt.item()) == hits)
        bf = n * (w / 8 + 1 / 8) + 8.0 * hits
        print(f"sel {sel:5.3f} hits {hits:10d}: scan {ms_s * 1e3:7.1f} us, positions {ms_p * 1e3:7.1f} us, fused call "
              f"{ms_f * 1e3:7.1f} us (two calls back to back {1e3 * ms_2:7.1f}, timed apart {1e3 * (ms_s + ms_p):7.1f}); fused {bf / ms_f / 1e6:7.1f} GB/s "
              f"{bf / ms_f / 1e6 / PEAK:.3f} of peak; list alone {(ms_f - ms_s) * 1e3:6.1f} us; equal={ok}")
        del pos, pos2


if __name__ == "__main__":
    main()
