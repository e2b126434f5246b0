#!/usr/bin/env python
"""The north-star table: shared scan (K bitvectors + K counts) and, optionally, the linear-layout
shared scan at EVERY bit width 1..32 and K in {1, 8, 64, 1024}, device-resident, one GPU.

    python tools/sweep_wk.py [--logn 30] [--ks 1,8,64,1024] [--ops shared,linear] [--out file.json]

n is 2^logn codes, reduced per K so that the output stays below --max-out-gib (K = 1024 at 2^30 codes
would be 128 GiB).  Bounds: K uniform-random ranges, lo uniform, width uniform in [0, 2*2^w/K)
(SURVEY.md section 8 d).  Times: CUDA events around up to 10 back-to-back launches, 2 warm-ups, best and
median of 5 pairs.  Prints one JSON row
per (op, w, K) and a per-K summary (min / median / max fraction of the measured HBM peak)."""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

PEAK = 6550.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, rounds=5, warm=2):
    """ms per launch: event pairs around enough back-to-back launches to fill about a millisecond (an
    event pair around ONE 50 us launch reads ~10 us too long); best and median of `rounds` pairs."""
    def once(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    reps = max(1, min(10, int(1.0 / max(once(1), 1e-3))))
    ts = [once(reps) for _ in range(rounds)]
    return min(ts), statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=30)
    ap.add_argument("--ks", default="1,8,64,1024")
    ap.add_argument("--ops", default="shared")
    ap.add_argument("--widths", default="1-32")
    ap.add_argument("--max-out-gib", type=float, default=16.0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lo_w, hi_w = (int(x) for x in a.widths.split("-")) if "-" in a.widths else (int(a.widths), int(a.widths))
    ks = [int(k) for k in a.ks.split(",")]
    ops = a.ops.split(",")
    n = 1 << a.logn
    max_out = int(a.max_out_gib * (1 << 30))
    # one output arena reused by every configuration (no allocator noise between rows)
    arena = torch.empty(max_out + (1 << 20), dtype=torch.uint8, device="cuda")
    rows = []
    for w in range(lo_w, hi_w + 1):
        m = (1 << w) - 1
        packed = api.generate_packed(n, w, seed=42 + w)
        for K in ks:
            n2 = n
            while K * (n2 // 8 + int(os.environ.get("SSS_STRIDE_PAD", "0")) + 160) > max_out:
                n2 //= 2
            g = torch.Generator().manual_seed(42 + K)
            lo = torch.randint(0, m + 1, (K,), generator=g)
            hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
            lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
            cnt = torch.empty(K, dtype=torch.int64, device="cuda")
            for op in ops:
                if op == "shared":
                    stride = (api.bitvec_bytes(n2) + 127) // 128 * 128 + int(os.environ.get("SSS_STRIDE_PAD", "0"))
                    out = arena[: K * stride].view(K, stride)
                    f = lambda: api.shared_scan(packed, n2, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
                else:
                    rows_ = 4 * ((n2 + 31) // 32)
                    out = arena[: rows_ * K].view(rows_, K)
                    f = lambda: api.shared_scan_linear(packed, n2, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
                best, med = timeit(f)
                b = n2 * (w / 8 + K / 8)
                row = dict(op=op, w=w, K=K, n=n2, ms=round(best, 4), ms_med=round(med, 4),
                           gcodes_per_s=round(n2 / best / 1e6, 2), gb_per_s=round(b / best / 1e6, 1),
                           frac_of_measured_peak=round(b / best / 1e6 / PEAK, 4),
                           frac_of_8TBps=round(b / best / 1e6 / 8000.0, 4))
                rows.append(row)
                print(json.dumps(row), flush=True)
        del packed
    summary = []
    for op in ops:
        for K in ks:
            fr = [r["frac_of_measured_peak"] for r in rows if r["op"] == op and r["K"] == K]
            if fr:
                summary.append(dict(op=op, K=K, min=min(fr), median=round(statistics.median(fr), 4), max=max(fr),
                                    argmin_w=[r["w"] for r in rows if r["op"] == op and r["K"] == K][fr.index(min(fr))]))
    for s in summary:
        print(json.dumps(s), flush=True)
    if a.out:
        json.dump({"peak_gb_per_s": PEAK, "rows": rows, "summary": summary}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
