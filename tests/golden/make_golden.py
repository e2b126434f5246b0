#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref/libref_wNN.so).

Run in the build container (needs /root/reference to have been compiled by oracle/build_ref.sh):
    python tests/golden/make_golden.py
Writes tests/golden/ref_wNN.npz.  The fixtures are committed; tests that use them need neither the
reference nor oracle/_ref at run time.

Per width w (only widths where the reference family is exact):
  codes   : n ragged (1003) uniform-random codes + a zero-free variant for the tail cases
  packed  : compress_9bit_input output (w <= 16); for w > 16 the reference has no compressor
            (uint16 input, src/simd_scan.hpp:46) so the stream comes from the oracle packer and the
            fixture says so (packed_by = "oracle"); it is still validated by the reference's
            decompress_128 reproducing `codes` from it.
  unpack  : decompress_128_unrolled (and decompress_256_avx2 for w <= 16)
  scan    : scan_128 / scan_128_unrolled (/ scan_256_unrolled, scan_unvectorized for w <= 16)
            whole output vectors + hits for several keys incl. 0 (tail padding matches)
  shared  : shared_scan_128_standard, _standard_unrolled, linear_standard, linear_simple (K=4)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
WIDTHS = [1, 2, 3, 5, 7, 8, 9, 11, 13, 16, 17, 20, 23, 26, 28, 32]
N = 1003


def s32(k: int) -> int:
    return k - (1 << 32) if k >= (1 << 31) else k


def main():
    for w in WIDTHS:
        assert B.ref_available(w), f"build oracle/_ref first (missing width {w})"
        rng = np.random.default_rng(1000 + w)
        hi = (1 << w) - 1
        codes = rng.integers(0, hi + 1, size=N, dtype=np.uint64).astype(np.uint32)
        if w <= 16:
            packed = B.ref_compress(codes.astype(np.uint16), w)
            packed_by = "reference"
        else:
            packed = B.pack(codes, w)
            packed_by = "oracle"
        d = {"w": w, "n": N, "codes": codes, "packed": packed, "packed_by": packed_by}
        d["unpack_128_unrolled"] = B.ref_decompress("128_unrolled", packed, N, w)
        assert (d["unpack_128_unrolled"].astype(np.uint32) == codes).all()
        if w <= 16:
            d["unpack_256_avx2"] = B.ref_decompress("256_avx2", packed, N, w)
            d["unpack_unvectorized"] = B.ref_decompress("unvectorized", packed, N, w)
        keys = sorted({0, int(codes[3]), int(codes[500]), hi})
        d["keys"] = np.array(keys, dtype=np.int64)
        variants = ["128", "128_unrolled"] + (["256_unrolled", "unvectorized"] if w <= 16 else [])
        for v in variants:
            outs, hits = [], []
            for k in keys:
                o, h = B.ref_scan(v, s32(k), packed, N, w)
                outs.append(o)
                hits.append(h)
            d[f"scan_{v}"] = np.stack(outs)
            d[f"hits_{v}"] = np.array(hits, dtype=np.int64)
        # incl. a duplicate key; the reference takes `int` keys, so w=32 codes >= 2^31 are passed as
        # the int with the same bit pattern (stored here as that int32)
        skeys = np.array([s32(int(codes[1])), 0, s32(int(codes[1])), s32(int(codes[999]))], dtype=np.int32)
        d["shared_keys"] = skeys
        d["shared_128_standard"] = B.ref_shared_scan("128_standard", skeys, packed, N, w)
        d["shared_128_standard_unrolled"] = B.ref_shared_scan("128_standard_unrolled", skeys, packed, N, w)
        d["linear_standard"] = B.ref_shared_scan_linear("linear_standard", skeys, packed, N, w)
        d["linear_simple"] = B.ref_shared_scan_linear("linear_simple", skeys, packed, N, w)
        np.savez_compressed(os.path.join(OUT, f"ref_w{w:02d}.npz"), **d)
        print("wrote", f"ref_w{w:02d}.npz")


if __name__ == "__main__":
    main()
