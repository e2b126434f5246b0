"""Position-list kernel: correctness against torch.nonzero on random bitvectors of several densities
(aligned and 4-byte-offset bitvector pointers, ragged n) and timing at 2^30 bits.
    python tools/pos_bench.py [--logn 30]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402


def random_bits(n_bytes, density, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = torch.zeros(n_bytes, dtype=torch.uint8, device="cuda")
    for b in range(8):
        out |= ((torch.rand(n_bytes, device="cuda", generator=g) < density).to(torch.uint8) << b)
    return out


def expected(bv, n):
    bits = ((bv[:, None] >> torch.arange(8, device="cuda", dtype=torch.uint8)[None, :]) & 1).reshape(-1)[:n]
    return torch.nonzero(bits).reshape(-1)


SIZES = (1, 31, 32, 33, 127, 4096 * 32 + 5, (1 << 18) + 3, (1 << 18) * 3 + 77, (1 << 22) + 12345)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=30)
    ap.add_argument("--peak", type=float, default=6550.7)
    ap.add_argument("--quick", action="store_true", help="timing only")
    args = ap.parse_args()
    bad = 0
    sizes = () if args.quick else SIZES
    for n in sizes:
        for density in (0.0, 0.0005, 0.02, 0.3, 0.6, 1.0):
            for off in (0, 4):
                buf = random_bits((n + 7) // 8 + 64, density, seed=n % 1000 + int(density * 1000))
                bv = buf[off:]
                exp = expected(bv, n)
                pos, num = api.positions(bv, n, base=7)
                k = int(num.item())
                ok = k == exp.numel() and bool((pos[:k] == exp + 7).all())
                if not ok:
                    bad += 1
                    print("MISMATCH n", n, "density", density, "off", off, "k", k, "exp", exp.numel())
    print("correctness:", "ok" if bad == 0 else f"{bad} mismatches")
    n = 1 << args.logn
    for density in (0.0, 0.001, 0.01, 0.05, 0.25, 0.5, 1.0):
        bv = random_bits(n // 8 + 64, density, seed=5)
        pos, num = api.positions(bv, n)
        hits = int(num.item())
        pos_buf = torch.empty(max(hits, 1), dtype=torch.int64, device="cuda")
        num_buf = torch.empty(1, dtype=torch.int64, device="cuda")
        del pos
        for _ in range(2):
            api.positions(bv, n, capacity=hits, out=pos_buf, num=num_buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            api.positions(bv, n, capacity=hits, out=pos_buf, num=num_buf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        b = n / 8 + 8.0 * hits
        asc = bool((pos_buf[1:hits] > pos_buf[: hits - 1]).all()) if hits > 1 else True
        print(f"density {density:6.3f}  hits {hits:11d}  {ms * 1000:8.1f} us  {b / ms / 1e6:7.1f} GB/s  "
              f"{b / ms / 1e6 / args.peak:.3f} of peak  ascending={asc}")
        del pos_buf, bv


if __name__ == "__main__":
    main()
