#!/usr/bin/env python
"""Developer micro-benchmark: per-width device-resident throughput of each kernel family.
Not the contract bench (that is bench.py); prints one line per (op, w[, K])."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

PEAK = 6550.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="scan,unpack,shared,linear")
    ap.add_argument("--widths", default="1-32")
    ap.add_argument("--logn", type=int, default=30)
    ap.add_argument("--ks", default="8,64")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lo_w, hi_w = (int(x) for x in a.widths.split("-")) if "-" in a.widths else (int(a.widths), int(a.widths))
    n = 1 << a.logn
    rows = []
    for op in a.ops.split(","):
        for w in range(lo_w, hi_w + 1):
            m = (1 << w) - 1
            packed = api.generate_packed(n, w, seed=42 + w)
            if op == "scan":
                out = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
                cnt = torch.empty(1, dtype=torch.int64, device="cuda")
                best, med = timeit(lambda: api.scan(packed, n, w, m // 4, m // 2, out=out, count=cnt))
                bpc = w / 8 + 1 / 8
                rows.append(dict(op=op, w=w, n=n, ms=best * 1e3, ms_med=med * 1e3, gcodes=n / best / 1e9,
                                 gbs=n * bpc / best / 1e9, frac=n * bpc / best / 1e9 / PEAK))
            elif op == "unpack":
                n2 = min(n, 1 << 28)
                out = torch.empty(n2, dtype=torch.int32, device="cuda")
                best, med = timeit(lambda: api.decompress(packed, n2, w, out=out))
                bpc = w / 8 + 4
                rows.append(dict(op=op, w=w, n=n2, ms=best * 1e3, ms_med=med * 1e3, gcodes=n2 / best / 1e9,
                                 gbs=n2 * bpc / best / 1e9, frac=n2 * bpc / best / 1e9 / PEAK))
                del out
            elif op in ("shared", "linear"):
                for K in (int(k) for k in a.ks.split(",")):
                    n2 = min(n, (1 << 33) // K)  # bound output to 1 GiB
                    g = torch.Generator().manual_seed(1)
                    lo = torch.randint(0, m + 1, (K,), generator=g)
                    hi = torch.clamp(lo + torch.randint(0, max(2, (m + 1) // K * 2), (K,), generator=g), max=m)
                    if os.environ.get("SSS_QB_EQ"):
                        hi = lo.clone()  # the reference's equality predicates
                    lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
                    cnt = torch.empty(K, dtype=torch.int64, device="cuda")
                    nocount = bool(os.environ.get("SSS_QB_NOCOUNT"))
                    if op == "shared":
                        out = torch.empty((K, api.bitvec_bytes(n2) // 16 * 16 + 16), dtype=torch.uint8, device="cuda")
                        if nocount:
                            from shared_simd_scan_b200._lib import lib, check
                            ctx = api.Context.current()
                            f = lambda: check(lib().sss_shared_scan(ctx.handle, packed.data_ptr(), n2, w, lo_d.data_ptr(), hi_d.data_ptr(), K, out.data_ptr(), out.stride(0), None, torch.cuda.current_stream().cuda_stream), "x")  # noqa: E731
                        else:
                            f = lambda: api.shared_scan(packed, n2, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
                    else:
                        out = torch.empty((4 * ((n2 + 31) // 32), K), dtype=torch.uint8, device="cuda")
                        f = lambda: api.shared_scan_linear(packed, n2, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
                    best, med = timeit(f, reps=3, warm=1)
                    bpc = w / 8 + K / 8
                    rows.append(dict(op=op, w=w, K=K, n=n2, ms=best * 1e3, ms_med=med * 1e3, gcodes=n2 / best / 1e9,
                                     gbs=n2 * bpc / best / 1e9, frac=n2 * bpc / best / 1e9 / PEAK))
                    del out
            print(json.dumps(rows[-1]), flush=True)
            del packed
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
