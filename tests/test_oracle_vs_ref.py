"""Oracle restatement == the real reference (oracle/_ref, built by oracle/build_ref.sh) on random
ragged inputs at every width where the reference family is exact.

Runs wherever oracle/_ref exists (this container and the GPU box, since the .so files travel);
skipped otherwise -- the committed goldens in test_oracle_golden.py still pin the oracle then.
"""
import numpy as np
import pytest

from oracle import binding as B

pytestmark = pytest.mark.skipif(not B.ref_available(9), reason="oracle/_ref not built")


def s32(k):
    k = int(k)
    return k - (1 << 32) if k >= (1 << 31) else k


@pytest.mark.parametrize("w", sorted(B.SSE_EXACT))
@pytest.mark.parametrize("n", [1, 31, 64, 4099])
def test_oracle_matches_reference(w, n):
    if not B.ref_available(w):
        pytest.skip("width not built")
    rng = np.random.default_rng(w * 100003 + n)
    codes = rng.integers(0, 1 << w, size=n, dtype=np.uint64).astype(np.uint32)
    packed = B.pack(codes, w)
    payload = (n * w + 7) // 8
    if w <= 16:
        rp = B.ref_compress(codes.astype(np.uint16), w)
        assert (rp[:payload] == packed[:payload]).all() and not rp[payload:].any()
    nb8, nb32 = (n + 7) // 8, 4 * ((n + 31) // 32)

    # unpack
    dec = ["128", "128_unrolled", "128_aligned"] + (["256", "256_avx2", "unvectorized"] if w <= 16 else [])
    mine = B.unpack(packed, n, w)
    for v in dec:
        assert (B.ref_decompress(v, packed, n, w) == mine).all(), v

    # single scan, a present key, key 0 (padding matches) and an absent/out-of-domain key
    for key in {int(codes[n // 2]), 0, (1 << w) - 1}:
        o8, h8 = B.scan_range(packed, n, w, key, key, granule=8)
        r, rh = B.ref_scan("128", s32(key), packed, n, w)
        assert (r[:nb8] == o8).all() and not r[nb8:].any() and rh == h8
        o32, h32 = B.scan_range(packed, n, w, key, key, granule=32)
        for v in ["128_unrolled"] + (["256_unrolled"] if w <= 16 else []):
            r, rh = B.ref_scan(v, s32(key), packed, n, w)
            assert (r[:nb32] == o32).all() and not r[nb32:].any() and rh == h32, v
        if w <= 16:
            r, rh = B.ref_scan("256", s32(key), packed, n, w)
            assert (r[:nb8] == o8).all() and rh == h8

    # shared scans, K = 5 with a duplicate
    keys = [int(codes[0]), int(codes[n - 1]), 0, int(codes[0]), (1 << w) - 1]
    skeys = [s32(k) for k in keys]
    o8, _ = B.shared_scan(packed, n, w, keys, keys, granule=8)
    o32, _ = B.shared_scan(packed, n, w, keys, keys, granule=32)
    for v in ["128_sequential", "128_threaded", "128_standard"] + (["256_standard"] if w <= 16 else []):
        r = B.ref_shared_scan(v, skeys, packed, n, w)
        assert (r[:, :nb8] == o8).all() and not r[:, nb8:].any(), v
    for v in ["128_sequential_unrolled", "128_standard_unrolled"]:
        r = B.ref_shared_scan(v, skeys, packed, n, w)
        assert (r[:, :nb32] == o32).all() and not r[:, nb32:].any(), v
    lin, _ = B.shared_scan_linear(packed, n, w, keys, keys, granule=8)
    r = B.ref_shared_scan_linear("linear_standard", skeys, packed, n, w)
    assert (r[: nb8 * 5] == lin.reshape(-1)).all() and not r[nb8 * 5:].any()
    # linear_simple only supports K in {1,2,4,...,1024} (src/simd_scan_shared_linear.cpp:66-81)
    lin4, _ = B.shared_scan_linear(packed, n, w, keys[:4], keys[:4], granule=8)
    r = B.ref_shared_scan_linear("linear_simple", skeys[:4], packed, n, w)
    assert (r[: nb8 * 4] == lin4.reshape(-1)).all() and not r[nb8 * 4:].any()


def test_linear_simple_k1024_duplicates():
    # K = 1024 at w = 9 forces duplicate keys (512 distinct codes): SURVEY.md section 8b
    w, n, K = 9, 515, 1024
    codes = B.generate_codes(n, w, seed=7)
    packed = B.pack(codes, w)
    keys = [(i * 7) % 512 for i in range(K)]
    lin, _ = B.shared_scan_linear(packed, n, w, keys, keys, granule=8)
    r = B.ref_shared_scan_linear("linear_simple", keys, packed, n, w)
    nb8 = (n + 7) // 8
    assert (r[: nb8 * K] == lin.reshape(-1)).all()
