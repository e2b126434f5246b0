#!/usr/bin/env python
"""Executable model of the word-parallel (SWAR) range scan used for small bit widths.

Validates, for every width, the exact 32-bit word operations the CUDA kernel performs on a 32-code
block of W words -- the guard-free field subtraction with a borrow chain, the MSB compare, and the
multiply-gather of the per-field result bits -- against a brute-force evaluation.
Run: python tools/swar_model.py
"""
import random

M32 = 0xFFFFFFFF


def tables(W):
    """Compile-time tables (mirrors swar.cuh): per word H mask, gather groups (mask, magic, count)."""
    H = [0] * W
    pos = [[] for _ in range(W)]
    for j in range(32):
        p = j * W + W - 1
        H[p >> 5] |= 1 << (p & 31)
        pos[p >> 5].append(p & 31)
    groups = []  # per word: list of (mask, magic, count), ascending bit order
    cap = max(W, 1)
    for i in range(W):
        gl = []
        ps = pos[i]
        for g0 in range(0, len(ps), cap):
            grp = ps[g0:g0 + cap]
            m = len(grp)
            mask = 0
            magic = 0
            for k, p in enumerate(grp):
                mask |= 1 << p
                e = (32 - m + k) - p
                assert e >= 0
                magic |= 1 << e
            gl.append((mask, magic & M32, m))
        groups.append(gl)
    return H, groups


def swar_block(x, W, lo, hi):
    """x: list of W 32-bit words (32 codes). Returns the 32-bit result word."""
    H, groups = tables(W)
    fm = (1 << W) - 1
    msb = 1 << (W - 1)
    d = hi - lo
    # replicated constants (host side in the real kernel)
    big = lambda v: sum(v << (j * W) for j in range(32))
    words = lambda b: [(b >> (32 * i)) & M32 for i in range(W)]
    LOL = words(big(lo & ~msb & fm))
    DH = words(big(d | msb))
    DRH = words(big(d & msb))
    lo_msb = (lo >> (W - 1)) & 1
    # step 1: T = (X | H) - LOL  (borrow chain)
    T, borrow = [], 0
    for i in range(W):
        a = x[i] | H[i]
        t = a - LOL[i] - borrow
        borrow = 1 if t < 0 else 0
        T.append(t & M32)
    assert borrow == 0
    # step 2: G = DH - (T & ~H)
    G, borrow = [], 0
    for i in range(W):
        t = DH[i] - (T[i] & ~H[i] & M32) - borrow
        borrow = 1 if t < 0 else 0
        G.append(t & M32)
    assert borrow == 0
    res = 0
    # V, R, gather (processed from the highest word/group down so code 0 ends at bit 0)
    for i in reversed(range(W)):
        L = ~H[i] & M32
        V = ((T[i] ^ x[i] ^ (0 if lo_msb else H[i])) | L) & M32
        R = ((DRH[i] & ~V) | (~(DRH[i] ^ V) & G[i])) & M32
        assert R & L == 0
        for (mask, magic, m) in reversed(groups[i]):
            prod = (R * magic) & M32  # no pre-mask: partial products of other result bits never reach the window
            res = ((res << m) | (prod >> (32 - m))) & M32
    return res


def pack_block(codes, W):
    b = sum(c << (j * W) for j, c in enumerate(codes))
    return [(b >> (32 * i)) & M32 for i in range(W)]


def main():
    rnd = random.Random(1)
    for W in range(2, 33):
        if W == 32:
            continue  # H/L split needs W < 32 for the model's int math; kernel uses the scalar path there
        fm = (1 << W) - 1
        for trial in range(300):
            lo = rnd.randint(0, fm)
            hi = rnd.randint(lo, fm)
            if trial % 7 == 0:
                hi = lo
            if trial % 11 == 0:
                lo, hi = 0, fm
            codes = [rnd.choice([lo, hi, max(lo - 1, 0), min(hi + 1, fm), rnd.randint(0, fm), 0, fm]) for _ in range(32)]
            x = pack_block(codes, W)
            got = swar_block(x, W, lo, hi)
            exp = sum(1 << j for j, c in enumerate(codes) if lo <= c <= hi)
            assert got == exp, (W, lo, hi, codes, hex(got), hex(exp))
        H, groups = tables(W)
        ng = sum(len(g) for g in groups)
        print(f"W={W:2d}: gather groups per block {ng:2d}, ops/block ~ {6 * W + 2 * ng:3d}  ({(6 * W + 2 * ng) / 32:.2f}/code)")
    print("model ok")


if __name__ == "__main__":
    main()
