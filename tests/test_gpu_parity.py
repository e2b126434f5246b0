"""GPU parity: the CUDA path, called through the C ABI, must be bit-exact with the CPU oracle
(and with the committed goldens of the unmodified reference) on the same seeded inputs.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import binding as B  # noqa: E402  (the checker, never the thing measured)

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_w*.npz")))
ALL_W = list(range(1, 33))


@pytest.fixture(scope="module")
def sss():
    import shared_simd_scan_b200 as m
    from shared_simd_scan_b200 import api
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m.Context.current()  # fails loudly if libsss_b200.so is missing or the device is unsuitable
    return api


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def column(n, w, seed):
    codes = B.generate_codes(n, w, seed=seed)
    return codes, B.pack(codes, w)


SIZES = [1, 31, 32, 33, 1000, 8192, 8192 * 3 + 5, 300_007]


@pytest.mark.parametrize("w", ALL_W)
def test_unpack_matches_oracle(sss, w):
    for n in SIZES:
        codes, packed = column(n, w, seed=100 + w)
        out = host(sss.decompress(dev(packed), n, w))
        assert (out == B.unpack(packed, n, w)).all(), (w, n)
        assert (out.astype(np.uint32) == codes).all()


@pytest.fixture(params=["direct", "ring"])
def scan_path(request, sss):
    """Short columns take the ring-less scan kernel, long ones the persistent TMA-ring kernel
    (csrc/scan.cu launch_mode).  The oracle comparisons use short columns, so they run twice: with the
    default context, and with one created under SSS_SCAN_DIRECT_TILES=0, which always takes the ring."""
    if request.param == "direct":
        yield sss
        return
    from shared_simd_scan_b200 import api
    dev_id = torch.cuda.current_device()
    old = api._contexts.get(dev_id)
    os.environ["SSS_SCAN_DIRECT_TILES"] = "0"
    try:
        api._contexts[dev_id] = api.Context(dev_id)
    finally:
        del os.environ["SSS_SCAN_DIRECT_TILES"]
    yield sss
    api._contexts[dev_id].close()
    api._contexts[dev_id] = old


@pytest.mark.parametrize("w", ALL_W)
def test_scan_matches_oracle(scan_path, w):
    sss = scan_path
    m = (1 << w) - 1
    for n in SIZES:
        codes, packed = column(n, w, seed=200 + w)
        d_packed = dev(packed)
        ranges = [(int(codes[n // 2]), int(codes[n // 2])), (0, 0), (m, m), (m // 4, m // 2 + 1), (0, m), (5, 4)]
        for lo, hi in ranges:
            bv, cnt = sss.scan(d_packed, n, w, lo, hi)
            exp, ecnt = B.scan_range(packed, n, w, lo, hi, granule=1)
            got = host(bv)[: 4 * ((n + 31) // 32)]
            assert (got[: len(exp)] == exp).all(), (w, n, lo, hi)
            assert not got[len(exp):].any(), "bits >= n must be zero"
            assert int(cnt.item()) == ecnt, (w, n, lo, hi)


@pytest.mark.parametrize("w", ALL_W)
def test_pack_matches_oracle(sss, w):
    for n in [1, 33, 1000, 70_001]:
        codes = B.generate_codes(n, w, seed=300 + w)
        exp = B.pack(codes, w)
        got = host(sss.compress(dev(codes.view(np.int32)), w))
        assert (got == exp).all(), (w, n)
        # device generator == host generator, packed
        gen = host(sss.generate_packed(n, w, seed=300 + w))
        assert (gen == exp).all(), (w, n)
        # first_index offsets the counter stream
        gen2 = host(sss.generate_packed(n - n // 2, w, seed=300 + w, first_index=n // 2))
        assert (B.unpack(gen2, n - n // 2, w).astype(np.uint32) == codes[n // 2:]).all()


@pytest.mark.parametrize("w", [1, 5, 9, 16, 17, 23, 31, 32])
@pytest.mark.parametrize("K", [1, 3, 4, 8, 64])
def test_shared_scan_matches_oracle(sss, w, K):
    m = (1 << w) - 1
    rng = np.random.default_rng(w * 1000 + K)
    for n in [31, 1000, 8192 * 2 + 77]:
        codes, packed = column(n, w, seed=400 + w)
        lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
        span = rng.integers(0, max(1, (m + 1) // max(K // 2, 1)) + 1, size=K, dtype=np.uint64)
        hi = np.minimum(lo + span, m)
        lo[0] = hi[0] = int(codes[0])          # an equality predicate that certainly hits
        if K > 2:
            lo[2], hi[2] = 7, 3                # an empty range
        exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
        bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
        got = host(bv)
        nb = (n + 7) // 8
        assert (got[:, :nb] == exp).all(), (w, K, n)
        assert not got[:, nb: 4 * ((n + 31) // 32)].any()
        assert (host(cnt).astype(np.uint64) == ecnt).all()

        lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
        lin, cnt2 = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
        lgot = host(lin)
        assert (lgot[:nb] == lexp).all(), (w, K, n)
        assert not lgot[nb:].any()
        assert (host(cnt2).astype(np.uint64) == lcnt).all()


def test_shared_scan_k1024_duplicates(sss):
    # config 4 shape (w = 9, K = 1024 => duplicate keys), small n
    w, n, K = 9, 5003, 1024
    codes, packed = column(n, w, seed=9)
    keys = np.array([(i * 7) % 512 for i in range(K)], dtype=np.uint32)
    lexp, lcnt = B.shared_scan_linear(packed, n, w, keys, keys, granule=1)
    lin, cnt = sss.shared_scan_linear(dev(packed), n, w, keys, keys)
    assert (host(lin)[: (n + 7) // 8] == lexp).all()
    assert (host(cnt).astype(np.uint64) == lcnt).all()
    exp, ecnt = B.shared_scan(packed, n, w, keys, keys, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, keys, keys)
    assert (host(bv)[:, : (n + 7) // 8] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()


@pytest.mark.parametrize("w", [3, 11, 32])
def test_positions(sss, w):
    for n in [1, 77, 40_000, 1_000_003]:
        codes, packed = column(n, w, seed=500 + w)
        m = (1 << w) - 1
        for lo, hi in [(0, m // 3), (int(codes[0]), int(codes[0])), (1, 0)]:
            bv, cnt = sss.scan(dev(packed), n, w, lo, hi)
            pos, num = sss.positions(bv, n, base=1 << 35)
            exp = B.positions(host(bv), n, base=1 << 35)
            k = int(num.item())
            assert k == len(exp) == int(cnt.item())
            assert (host(pos)[:k].astype(np.uint64) == exp).all()
        # capacity smaller than the number of matches: first `capacity` positions, full count
        bv, cnt = sss.scan(dev(packed), n, w, 0, m)
        pos, num = sss.positions(bv, n, capacity=min(n, 10))
        assert int(num.item()) == n and (host(pos)[: min(n, 10)] == np.arange(min(n, 10))).all()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_host_api_reproduces_reference_goldens(sss, path):
    """The host-buffer C-ABI calls (what include/simd_scan.hpp wraps) against vectors produced by the
    unmodified reference, including each variant's tail granule."""
    g = np.load(path)
    w, n = int(g["w"]), int(g["n"])
    packed = np.ascontiguousarray(g["packed"])
    u32 = lambda k: int(k) & 0xFFFFFFFF  # noqa: E731
    assert (sss.host_decompress(packed, n, w).astype(np.uint32) == g["codes"]).all()
    if g["packed_by"] == "reference":
        assert (sss.host_compress(g["codes"], w)[: len(packed)] == packed[: sss.packed_bytes(n, w)]).all()
    for i, key in enumerate(g["keys"]):
        out, hits = sss.host_scan(packed, n, w, u32(key), tail_granule=8)
        assert (out == g["scan_128"][i]).all() and hits == int(g["hits_128"][i])
        out, hits = sss.host_scan(packed, n, w, u32(key), tail_granule=32)
        assert (out == g["scan_128_unrolled"][i]).all() and hits == int(g["hits_128_unrolled"][i])
    skeys = [u32(k) for k in g["shared_keys"]]
    outs, _ = sss.host_shared_scan(packed, n, w, skeys, tail_granule=8)
    assert (outs == g["shared_128_standard"]).all()
    outs, _ = sss.host_shared_scan(packed, n, w, skeys, tail_granule=32)
    assert (outs == g["shared_128_standard_unrolled"]).all()
    lin, _ = sss.host_shared_scan_linear(packed, n, w, skeys, tail_granule=8)
    assert (lin == g["linear_standard"]).all() and (lin == g["linear_simple"]).all()


def test_host_api_chunked_pipeline(sss):
    """n large enough for several staging chunks (64 MiB each): equals the device-resident path."""
    w, n = 13, (1 << 26) + 12345
    d_packed = sss.generate_packed(n, w, seed=77)
    torch.cuda.synchronize()
    packed = host(d_packed)
    lo, hi = 100, 4000
    bv, cnt = sss.scan(d_packed, n, w, lo, hi)
    out, hits = sss.host_scan(packed, n, w, lo, hi, tail_granule=1)
    nb = (n + 7) // 8
    assert (out[:nb] == host(bv)[:nb]).all() and hits == int(cnt.item())
    # and a sample of it against the oracle
    m = 1 << 20
    exp, _ = B.scan_range(packed, m, w, lo, hi, granule=1)
    assert (out[: m // 8] == exp).all()
    dec = sss.host_decompress(packed, n, w)
    assert (dec == host(sss.decompress(d_packed, n, w))).all()


@pytest.mark.parametrize("w", [1, 4, 11, 17, 32])
def test_full_size_properties(sss, w):
    """BASELINE sizes (2^30 codes) through size-independent properties: counts add up over a
    partition of the domain, bitvectors of the parts are disjoint and OR to the whole, count ==
    popcount, unpack -> pack round trip, and a window of the result equals the oracle."""
    n = 1 << 30
    m = (1 << w) - 1
    d_packed = sss.generate_packed(n, w, seed=42 + w)
    mid = m // 3
    bv_a, c_a = sss.scan(d_packed, n, w, 0, mid)
    bv_b, c_b = sss.scan(d_packed, n, w, mid + 1, m)
    bv_all, c_all = sss.scan(d_packed, n, w, 0, m)
    assert int(c_all.item()) == n and int(c_a.item()) + int(c_b.item()) == n
    nb = n // 8
    a64, b64 = bv_a[:nb].view(torch.int64), bv_b[:nb].view(torch.int64)
    assert not (a64 & b64).any() and ((a64 | b64) == -1).all()
    # popcount(bitvec) == count (per-byte LUT on the device)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device="cuda")
    pop = 0
    for part in bv_a[:nb].split(1 << 27):
        pop += int(lut[part.long()].sum().item())
    assert pop == int(c_a.item())
    # window vs oracle
    win = 1 << 16
    first = (n // 2 // 1024) * 1024
    byte0 = first * w // 8
    host_packed = np.zeros(B.compressed_buffer_size(w, win), dtype=np.uint8)
    chunk = host(d_packed[byte0: byte0 + (win * w + 7) // 8])
    host_packed[: len(chunk)] = chunk
    exp, _ = B.scan_range(host_packed, win, w, 0, mid, granule=1)
    assert (host(bv_a[first // 8: first // 8 + win // 8]) == exp).all()
    del bv_a, bv_b, bv_all
    # unpack -> pack round trip on 2^28 codes
    n2 = 1 << 28
    dec = sss.decompress(d_packed, n2, w)
    re = sss.compress(dec, w)
    payload = n2 * w // 8
    assert torch.equal(re[:payload], d_packed[:payload])


def test_range_shards_concatenate_to_the_whole(sss):
    """The multi-GPU contract on one GPU: shards cut at multiples of 128 codes, scanned
    independently from sub-buffers of the column, concatenate to the single-GPU result; counts add;
    positions rebased by the shard base are globally ascending."""
    from shared_simd_scan_b200 import dist as sdist
    w, n, world = 23, 1_000_003, 4
    codes, packed = column(n, w, seed=23)
    d_packed = dev(packed)
    m = (1 << w) - 1
    lo, hi = m // 7, m // 3
    whole, cnt = sss.scan(d_packed, n, w, lo, hi)
    pieces, total, allpos = [], 0, []
    for r in range(world):
        first, count = sdist.shard_range(n, world, r)
        assert first % 128 == 0
        sub = torch.zeros(sss.packed_bytes(count, w), dtype=torch.uint8, device="cuda")
        nb = (count * w + 7) // 8
        sub[:nb] = d_packed[first * w // 8: first * w // 8 + nb]
        if (count * w) % 8:  # clear the bits of the next shard's first code in the last byte
            sub[nb - 1] &= (1 << ((count * w) % 8)) - 1
        bv, c = sss.scan(sub, count, w, lo, hi)
        pieces.append(host(bv)[: (count + 7) // 8])
        total += int(c.item())
        pos, num = sss.positions(bv, count, base=first)
        allpos.append(host(pos)[: int(num.item())])
    assert total == int(cnt.item())
    assert (np.concatenate(pieces)[: (n + 7) // 8] == host(whole)[: (n + 7) // 8]).all()
    pos, num = sss.positions(whole, n)
    assert (np.concatenate(allpos) == host(pos)[: int(num.item())]).all()


@pytest.mark.parametrize("w", [14, 17, 25, 32])
def test_shared_scan_clustered_bounds_and_codes(sss, w):
    """Many predicates and most codes inside a handful of locator buckets: exercises the path for
    buckets that hold several breakpoints (rare on uniform data, so forced here)."""
    n, K = 20_011, 96
    rng = np.random.default_rng(w)
    m = (1 << w) - 1
    center = m // 3
    codes = (center + rng.integers(-300, 300, size=n)).astype(np.int64)
    codes[::7] = rng.integers(0, m + 1, size=len(codes[::7]))       # some far-away codes too
    codes = np.clip(codes, 0, m).astype(np.uint32)
    packed = B.pack(codes, w)
    lo = np.clip(center + rng.integers(-320, 320, size=K), 0, m).astype(np.uint64)
    hi = np.minimum(lo + rng.integers(0, 40, size=K).astype(np.uint64), m)
    lo[5], hi[5] = 0, m          # everything
    lo[6], hi[6] = m, m          # the top code only
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt2 = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
    assert (host(lin)[:nb] == lexp).all()
    assert (host(cnt2).astype(np.uint64) == lcnt).all()


@pytest.mark.parametrize("w", list(range(12, 33)))
@pytest.mark.parametrize("K", [1, 5, 8])
def test_few_predicates_clustered(sss, w, K):
    """K <= 8 at every width around the mask-carrying bucket table (14..28 bit; 12, 13 and 29, 30 take
    the neighbouring kernels): codes and bounds crowded into a few buckets so that buckets with
    several breakpoints, equality predicates, ties on a threshold and the bucket's first / last code
    all occur; ragged column, both layouts, counts."""
    n = 50_021
    rng = np.random.default_rng(1000 * w + K)
    m = (1 << w) - 1
    span = 1 << max(w - 13, 1)                          # codes per 13-bit bucket
    center = (m // 3) // span * span                     # a bucket boundary
    codes = (center + rng.integers(-3 * span, 3 * span, size=n)).astype(np.int64)
    codes[::5] = rng.integers(0, m + 1, size=len(codes[::5]))
    codes[1::11] = center                                # the bucket's first code
    codes[2::11] = center - 1                            # the previous bucket's last code
    codes = np.clip(codes, 0, m).astype(np.uint32)
    packed = B.pack(codes, w)
    lo = np.clip(center + rng.integers(-2 * span, 2 * span, size=K), 0, m).astype(np.uint64)
    hi = np.minimum(lo + rng.integers(0, span, size=K).astype(np.uint64), m)
    lo[0], hi[0] = center, center                        # equality on a bucket boundary
    if K > 2:
        lo[1], hi[1] = center - 1, center + span - 1     # ends exactly on a bucket's last code
        lo[2], hi[2] = 0, m                              # everything
    if K > 4:
        lo[3], hi[3] = 7, 3                              # empty
        lo[4], hi[4] = m, m                              # the top code only
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt2 = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
    assert (host(lin)[:nb] == lexp).all()
    assert (host(cnt2).astype(np.uint64) == lcnt).all()


@pytest.mark.parametrize("K", [1, 5, 8])
@pytest.mark.parametrize("w", [27, 28, 29, 30, 31, 32])
def test_few_predicates_wide_codes_at_the_bounds(sss, w, K):
    """K <= 8 around the seam between the mask-carrying bucket table (narrow.cu, up to 28 bit: 15 threshold bits
    in the table word) and the thread-per-block kernel (29-32 bit: 16-19 bits below the bucket bits).  Most codes
    sit within +-40 of a bound -- equal to a threshold in all but their lowest bits, on both sides -- the rest is
    uniform; equality predicates, adjacent bounds, the domain's ends; ragged column, both layouts, counts.  (Written
    for an extension of the table to 29-32 bit with tie patching, which was bit-exact but no faster:
    profiles/r02_few_wide_experiment.txt.)"""
    n = 40_009
    rng = np.random.default_rng(5000 * w + K)
    m = (1 << w) - 1
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, 1 << (w - 6), size=K, dtype=np.uint64), m)
    lo[0] = hi[0] = lo[0] | 0xF                      # an equality predicate whose low four bits are all ones
    if K > 2:
        lo[1], hi[1] = 0, 16                         # starts at the domain's first code
        lo[2], hi[2] = m - 17, m                     # ends at its last
    if K > 4:
        lo[3], hi[3] = min(int(hi[0]) + 1, m), min(int(hi[0]) + 3, m)   # adjacent to predicate 0: thresholds 1 and 4 codes apart
        lo[4] = hi[4] = (int(lo[4]) >> 4 << 4)       # an equality predicate whose low four bits are zero
    bounds = np.concatenate([lo, hi]).astype(np.int64)
    codes = bounds[rng.integers(0, len(bounds), size=n)] + rng.integers(-40, 41, size=n)
    codes[::7] = rng.integers(0, m + 1, size=len(codes[::7]))
    codes = np.clip(codes, 0, m).astype(np.uint32)
    packed = B.pack(codes, w)
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt2 = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
    assert (host(lin)[:nb] == lexp).all()
    assert (host(cnt2).astype(np.uint64) == lcnt).all()


def test_predicate_groups_match_single_scans(sss):
    """K = 1024 at 23 bit is evaluated in predicate groups (the merged row table of all 1024 does not
    fit in shared memory): every bitvector and count must equal the oracle's, duplicates and empty
    predicates across group boundaries included."""
    w, K, n = 23, 1024, 30_007
    rng = np.random.default_rng(77)
    m = (1 << w) - 1
    codes, packed = column(n, w, seed=4242)
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, (m + 1) // 16, size=K, dtype=np.uint64), m)
    lo[511], hi[511] = lo[512], hi[512]                  # duplicates across a group boundary
    lo[255], hi[255] = 9, 2                              # an empty one at the end of a group
    lo[768], hi[768] = 0, m                              # everything, first of a group
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()


@pytest.mark.parametrize("w,K", [(9, 1024), (17, 1000), (12, 600), (32, 515), (5, 1024)])
def test_many_predicates(sss, w, K):
    """K > 512 takes the warp-per-block linear kernel and the histogram counts at full width."""
    n = 9_973
    rng = np.random.default_rng(w * 7 + K)
    m = (1 << w) - 1
    codes, packed = column(n, w, seed=600 + w)
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, max(2, (m + 1) // 64), size=K, dtype=np.uint64), m)
    hi[::5] = lo[::5]                       # equality predicates among them
    lo[3], hi[3] = 9, 2                      # an empty one
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(lin)[:nb] == lexp).all()
    assert not host(lin)[nb:].any()
    assert (host(cnt).astype(np.uint64) == lcnt).all()
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()


@pytest.mark.parametrize("w", [2, 3, 4, 6, 8, 13, 20, 24, 28])
@pytest.mark.parametrize("K", [2, 7, 16, 100, 128, 512])
def test_shared_scan_store_paths_and_row_tables(sss, w, K):
    """Narrow codes with few predicates (the group-table kernel: W <= 6, K <= 8), widths with rotated
    block loads (8, 24), 64-bit (2 mod 4) and 128-bit (4 mod 8) block loads, and
    predicate counts that exercise the single-store rows (K = 2), 16-byte row pieces (K = 16), ragged
    columns (K = 100) and partially replicated row tables (K = 128, 512), over several tiles."""
    n = 40_003
    rng = np.random.default_rng(w * 97 + K)
    m = (1 << w) - 1
    codes, packed = column(n, w, seed=700 + w)
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, max(2, (m + 1) // 8), size=K, dtype=np.uint64), m)
    hi[::3] = lo[::3]
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt2 = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
    assert (host(lin)[:nb] == lexp).all()
    assert not host(lin)[nb:].any()
    assert (host(cnt2).astype(np.uint64) == lcnt).all()


def test_shared_scan_wide_kernel_unaligned_stride(sss):
    """K > 512 with a bitvector stride that is not a multiple of 32 bytes: the warp-per-block kernel
    must fall back from its 256-bit stores to word stores."""
    w, K, n = 24, 640, 33_000
    rng = np.random.default_rng(5)
    m = (1 << w) - 1
    codes, packed = column(n, w, seed=811)
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, 200, size=K, dtype=np.uint64), m)
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    stride = 4 * ((n + 31) // 32) + 4                      # whole words, but 4 mod 32 bytes
    out = torch.zeros((K, stride), dtype=torch.uint8, device="cuda")
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi, out=out)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()


@pytest.mark.parametrize("w,K", [(17, 64), (9, 64), (23, 8)])
def test_shared_scan_full_size_properties(sss, w, K):
    """2^28 codes: every shared-scan bitvector equals the single scan of its predicate (itself
    oracle-checked above), counts equal, and the linear layout holds the same bits."""
    n = 1 << 28
    m = (1 << w) - 1
    d_packed = sss.generate_packed(n, w, seed=42 + w)
    g = torch.Generator().manual_seed(w * 131 + K)
    lo = torch.randint(0, m + 1, (K,), generator=g)
    hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
    hi[1] = lo[1]                      # an equality predicate
    lo_l, hi_l = lo.tolist(), hi.tolist()
    bvs, counts = sss.shared_scan(d_packed, n, w, lo_l, hi_l)
    nb = n // 8
    for k in sorted({0, 1, K // 2, K - 1}):
        bv, c = sss.scan(d_packed, n, w, lo_l[k], hi_l[k])
        assert torch.equal(bvs[k, :nb], bv[:nb]), (w, K, k)
        assert int(counts[k].item()) == int(c.item())
    if K * (n // 8) <= (1 << 31):
        lin, lcounts = sss.shared_scan_linear(d_packed, n, w, lo_l, hi_l)
        assert torch.equal(lcounts, counts)
        for k in (1, K - 1):
            assert torch.equal(lin[:nb, k].contiguous(), bvs[k, :nb])


def test_c_abi_argument_errors(sss):
    """The boundary reports bad arguments with status codes instead of faulting."""
    import ctypes as C
    from shared_simd_scan_b200 import _lib
    L = _lib.lib()
    ctx = sss.Context.current().handle
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    out = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    st = L.sss_scan_range(ctx, buf.data_ptr(), 100, 0, 1, 2, out.data_ptr(), None, None)
    assert st == -1                                      # w = 0
    st = L.sss_scan_range(ctx, buf.data_ptr(), 100, 33, 1, 2, out.data_ptr(), None, None)
    assert st == -1                                      # w = 33
    st = L.sss_scan_range(ctx, buf.data_ptr() + 4, 100, 9, 1, 2, out.data_ptr(), None, None)
    assert st == -1                                      # packed not 16-byte aligned
    st = L.sss_scan_range(ctx, None, 100, 9, 1, 2, out.data_ptr(), None, None)
    assert st == -1
    lo = torch.zeros(2048, dtype=torch.int32, device="cuda")
    st = L.sss_shared_scan(ctx, buf.data_ptr(), 100, 9, lo.data_ptr(), lo.data_ptr(), 1025, out.data_ptr(), 64, None, None)
    assert st == -1                                      # K > SSS_MAX_PREDICATES
    st = L.sss_shared_scan(ctx, buf.data_ptr(), 100, 9, lo.data_ptr(), lo.data_ptr(), 4, out.data_ptr(), 6, None, None)
    assert st == -1                                      # stride not a multiple of 4 / too small
    h = C.c_uint64(0)
    st = L.sss_host_scan_range(ctx, buf.cpu().numpy().ctypes.data, 100, 9, 1, 2, 5, out.cpu().numpy().ctypes.data, C.byref(h))
    assert st == -1                                      # unknown tail mode
    assert b"invalid" in L.sss_error_string(-1)
    # n == 0 is legal and touches nothing
    cnt = torch.full((1,), 7, dtype=torch.int64, device="cuda")
    assert L.sss_scan_range(ctx, buf.data_ptr(), 0, 9, 1, 2, out.data_ptr(), cnt.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0 and not out.any()


@pytest.mark.parametrize("op", ["and", "or", "andnot", "xor"])
def test_bitvec_combine(sss, op):
    for n in [1, 37, 128, 100_003, (1 << 22) + 5]:
        w = 12
        codes, packed = column(n, w, seed=700 + n % 97)
        d = dev(packed)
        a, _ = sss.scan(d, n, w, 100, 3000)
        b, _ = sss.scan(d, n, w, 2000, 4095)
        out, cnt = sss.bitvec_combine(op, a, b, n)
        exp, ehits = B.bitvec_combine(op, host(a), host(b), n)
        assert (host(out)[: len(exp)] == exp).all() and int(cnt.item()) == ehits, (op, n)
        # in place
        a2 = a.clone()
        sss.bitvec_combine(op, a2, b, n, out=a2)
        assert torch.equal(a2[: len(exp)], out[: len(exp)])


@pytest.mark.parametrize("w", [1, 9, 17, 32])
def test_late_materialisation(sss, w):
    """scan one column -> position list -> fetch the matching rows of ANOTHER packed column."""
    n = 200_003
    codes_a, packed_a = column(n, 11, seed=801)
    codes_b, packed_b = column(n, w, seed=802 + w)
    bv, cnt = sss.scan(dev(packed_a), n, 11, 17, 300)
    pos, num = sss.positions(bv, n)
    k = int(num.item())
    got = host(sss.decompress_at(dev(packed_b), w, pos, num=k))
    exp = B.unpack_at(packed_b, w, host(pos)[:k].astype(np.uint64))
    assert (got == exp).all()
    assert (got.astype(np.uint32) == codes_b[(codes_a >= 17) & (codes_a <= 300)]).all()


@pytest.mark.gpu
def test_calibration_kernels_move_the_bytes(sss):
    """sss_calibrate: copy reproduces the source, write fills the target, mix writes bytes/ratio."""
    nbytes = (1 << 22) + 4096
    a = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda")
    b = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    sss.calibrate("copy", a, b, nbytes)
    assert torch.equal(a, b)
    b.zero_()
    sss.calibrate("write", None, b, nbytes)
    assert int((b.view(torch.int32).view(-1, 4)[:, 1] == 1).sum().item()) == nbytes // 16
    sss.calibrate("read", a, None, nbytes)
    b.zero_()
    sss.calibrate("mix", a, b, nbytes, 4)
    torch.cuda.synchronize()
    assert not b[nbytes // 4 + 64:].any()          # nothing beyond bytes/ratio is written


# ---- the benchmarked grid, cell by cell ---------------------------------------------------------------
# bench.py's shared_sweep publishes a number for every (op, w, K) with w = 1..32 and K in {1, 8, 64, 1024}.
# Every width is its own template<int W> instantiation with width-class-specific loads and rotations, and
# every K class takes a different kernel (scan kernel / narrow + few-predicate kernels / thread-per-block /
# predicate groups / warp-per-block), so every cell gets its own oracle comparison: uniform ranges drawn
# the way the bench draws them, plus one equality predicate that certainly hits, one empty range, one
# full-domain range; ragged n; both layouts; counts.
@pytest.mark.parametrize("K", [1, 8, 64, 1024])
@pytest.mark.parametrize("w", ALL_W)
def test_benchmarked_grid_matches_oracle(sss, w, K):
    m = (1 << w) - 1
    rng = np.random.default_rng(50_000 + 100 * w + K)
    n = 8192 * 2 + 77
    codes, packed = column(n, w, seed=700 + w)
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, max(2, 2 * (m + 1) // K), size=K, dtype=np.uint64), m)
    if K == 1:
        lo[0] = hi[0] = int(codes[n // 2])                 # the reference's own semantics: equality
    else:
        lo[1] = hi[1] = int(codes[n // 2])                 # an equality predicate that certainly hits
        lo[K // 2], hi[K // 2] = min(7, m), min(3, m - 1) if m > 0 else 0   # an empty range (lo > hi)
        if lo[K // 2] <= hi[K // 2]:
            lo[K // 2], hi[K // 2] = 1, 0
        lo[K - 1], hi[K - 1] = 0, m                        # everything
    d_packed = dev(packed)
    nb = (n + 7) // 8
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(d_packed, n, w, lo, hi)
    got = host(bv)
    assert (got[:, :nb] == exp).all(), (w, K)
    assert not got[:, nb: 4 * ((n + 31) // 32)].any(), "bits >= n must be zero"
    assert (host(cnt).astype(np.uint64) == ecnt).all(), (w, K)
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt2 = sss.shared_scan_linear(d_packed, n, w, lo, hi)
    lgot = host(lin)
    assert (lgot[:nb] == lexp).all(), (w, K)
    assert not lgot[nb:].any()
    assert (host(cnt2).astype(np.uint64) == lcnt).all(), (w, K)


@pytest.mark.parametrize("w", [2, 9, 17, 32])
def test_scan_positions_one_call(sss, w):
    """sss_scan_range_positions: bitvector + count + ascending positions (+ base) from one call."""
    m = (1 << w) - 1
    for n in [1, 77, 65_536 + 3, 1_000_003]:
        codes, packed = column(n, w, seed=900 + w)
        for lo, hi in [(0, m // 3), (int(codes[0]), int(codes[0])), (1, 0), (0, m)]:
            bv, pos, cnt = sss.scan_positions(dev(packed), n, w, lo, hi, base=1 << 40)
            exp, ecnt = B.scan_range(packed, n, w, lo, hi, granule=1)
            assert (host(bv)[: len(exp)] == exp).all() and int(cnt.item()) == ecnt
            epos = B.positions(exp, n, base=1 << 40)
            assert (host(pos)[:ecnt].astype(np.uint64) == epos).all(), (w, n, lo, hi)


@pytest.mark.parametrize("w", ALL_W)
def test_single_predicate_shared_scan_both_kernels(scan_path, w):
    """K = 1: the shared scan IS the scan kernel with its bounds read on the device -- the ring-less kernel for
    1-2 bit codes and short columns (scan_direct_dev_kernel), the persistent ring kernel otherwise
    (scan_range_dev_kernel; forced by the `ring` variant of the fixture).  Both layouts, the count, an empty
    and an out-of-domain predicate, ragged sizes."""
    sss = scan_path
    m = (1 << w) - 1
    for n in [1, 33, 8192 * 3 + 5, 300_007]:
        codes, packed = column(n, w, seed=400 + w)
        d_packed = dev(packed)
        nb = (n + 7) // 8
        for lo, hi in [(m // 4, m // 2), (int(codes[n // 2]), int(codes[n // 2])), (0, m), (min(3, m), min(3, m) - 1)]:
            exp, ecnt = B.shared_scan(packed, n, w, [lo], [hi], granule=1)
            got, cnt = sss.shared_scan(d_packed, n, w, [lo], [hi])
            assert (host(got)[:, :nb] == exp).all(), (w, n, lo, hi)
            assert not host(got)[:, nb: 4 * ((n + 31) // 32)].any()
            assert (host(cnt).astype(np.uint64) == ecnt).all(), (w, n, lo, hi)
            lin, cnt2 = sss.shared_scan_linear(d_packed, n, w, [lo], [hi])
            assert (host(lin).reshape(-1)[:nb] == exp[0]).all() and (host(cnt2).astype(np.uint64) == ecnt).all()


@pytest.mark.parametrize("w", ALL_W)
def test_scan_positions_fused_counts_every_width(scan_path, w):
    """sss_scan_range_positions lets the SCAN fill the position list's segment counts (scan.cu kSeg: one
    warp-aggregated atomic per warp and 1024-word segment) instead of a count pass over the bitvector.
    Every width has its own tile geometry (256 .. 8192 result words per tile, 1 .. 8 rounds), so all 32 are
    checked against the oracle, on both scan kernels, at sizes around the segment and tile borders."""
    sss = scan_path
    m = (1 << w) - 1
    for n in [1, 32767, 32768 * 8 + 1, 300_007]:
        codes, packed = column(n, w, seed=950 + w)
        d_packed = dev(packed)
        for lo, hi in [(m // 4, m // 2), (int(codes[n // 3]), int(codes[n // 3])), (0, m)]:
            bv, pos, cnt = sss.scan_positions(d_packed, n, w, lo, hi, base=7)
            exp, ecnt = B.scan_range(packed, n, w, lo, hi, granule=1)
            assert (host(bv)[: len(exp)] == exp).all() and int(cnt.item()) == ecnt, (w, n, lo, hi)
            assert (host(pos)[:ecnt].astype(np.uint64) == B.positions(exp, n, base=7)).all(), (w, n, lo, hi)


@pytest.mark.parametrize("w", [2, 3, 4, 6, 9, 11, 13, 16, 17, 23, 32])
def test_scan_positions_fused_equals_two_calls_long_column(sss, w):
    """Long columns (the persistent ring kernel with thousands of tiles drawn from the ticket counter): the
    fused call must give exactly what sss_scan_range followed by sss_positions gives."""
    m = (1 << w) - 1
    n = (1 << (29 if w <= 4 else 27)) + 12_345
    d_packed = sss.generate_packed(n, w, seed=77 + w)
    for lo, hi in [(m // 4, m // 2), (1 & m, 1 & m)]:
        bv0, cnt0 = sss.scan(d_packed, n, w, lo, hi)
        pos0, num0 = sss.positions(bv0, n, base=3, capacity=int(cnt0.item()))
        bv, pos, cnt = sss.scan_positions(d_packed, n, w, lo, hi, base=3, capacity=int(cnt0.item()))
        assert int(cnt.item()) == int(cnt0.item()) == int(num0.item())
        nb = 4 * ((n + 31) // 32)  # (the buffers' padding beyond the result words is never written)
        assert torch.equal(bv[:nb], bv0[:nb])
        assert torch.equal(pos[: int(cnt0.item())], pos0[: int(cnt0.item())]), (w, lo, hi)
        del bv0, pos0, bv, pos


@pytest.mark.parametrize("w", [1, 2, 5, 9, 11, 17, 32])
def test_programmatic_launch_keeps_stream_order(sss, w):
    """The ring-less scan kernel is launched with programmatic stream serialization: its CTAs may be scheduled
    while the previous kernel of the stream is still running and must wait (griddepcontrol.wait) before they
    touch memory.  Here the previous kernel PRODUCES the scan's input -- the compressor writes a new column into
    the same buffer before every scan, nothing synchronises in between -- so a scan that started early would
    count the previous column."""
    n = (1 << 20) + 77
    m = (1 << w) - 1
    lo, hi = m // 4, m // 2
    g = torch.Generator(device="cuda").manual_seed(w)
    packed = torch.zeros(sss.packed_bytes(n, w), dtype=torch.uint8, device="cuda")
    bv = torch.empty(sss.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    columns = [torch.randint(0, m + 1, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32) for _ in range(24)]
    expect = [int(((c.to(torch.int64) & 0xFFFFFFFF >= lo) & (c.to(torch.int64) & 0xFFFFFFFF <= hi)).sum().item()) for c in columns]
    torch.cuda.synchronize()
    got = []
    for c in columns:
        sss.compress(c, w, out=packed)
        sss.scan(packed, n, w, lo, hi, out=bv, count=cnt)
        sss.scan(packed, n, w, lo, hi, out=bv, count=cnt)   # scan after scan: the same count again
        got.append(cnt.clone())
    torch.cuda.synchronize()
    assert [int(x.item()) for x in got] == expect


def test_positions_many_tiles_and_reuse(sss):
    """The single-pass position list across thousands of look-back tiles, dense and sparse, called
    repeatedly on one context (epoch-tagged tile states: no clearing between launches) and with
    different sizes back to back."""
    w = 7
    m = (1 << w) - 1
    n_big = (1 << 27) + 12_345
    d_packed = sss.generate_packed(n_big, w, seed=31)
    for rep, (n, lo, hi) in enumerate([(n_big, 0, m // 2), (1 << 20, 0, 0), (n_big, 5, 5), (3_000_001, 0, m),
                                        (n_big, 0, m // 2)]):
        bv, cnt = sss.scan(d_packed, n, w, lo, hi)
        pos, num = sss.positions(bv, n, base=rep, capacity=int(cnt.item()))
        k = int(num.item())
        assert k == int(cnt.item())
        p = pos[:k]
        if k > 1:
            assert bool((p[1:] > p[:-1]).all()), "positions must be strictly ascending"
        # every listed position has its bit set, and there are exactly popcount of them
        idx = p - rep
        bits = (bv[(idx >> 3)].to(torch.int64) >> (idx & 7)) & 1
        assert bool(bits.all())
    # a window against the oracle
    n = 1 << 18
    packed = host(d_packed[: n * w // 8 + 300])
    exp, _ = B.scan_range(packed, n, w, 0, m // 2, granule=1)
    bv, cnt = sss.scan(d_packed, n, w, 0, m // 2)
    pos, num = sss.positions(bv, n)
    assert (host(pos)[: int(num.item())].astype(np.uint64) == B.positions(exp, n)).all()


@pytest.mark.parametrize("density", [0.0, 0.0005, 0.02, 0.3, 0.6, 1.0])
def test_positions_every_density_and_unaligned_bitvector(sss, density):
    """Position list of random bitvectors from empty to full (the staged path, the dense direct path and
    rounds that switch between them), ragged sizes around the 1024-word segment, and a bitvector whose
    address is only 4-byte aligned (sss_positions accepts that; the kernel then loads word by word)."""
    rng = np.random.default_rng(int(density * 10_000) + 3)
    for n in [1, 31, 33, 32767, 32768, 32769, 4096 * 32 + 5, (1 << 18) * 3 + 77]:
        nbytes = (n + 7) // 8
        bits = rng.random(nbytes * 8) < density
        bv_h = np.packbits(bits.reshape(-1, 8)[:, ::-1], axis=1).reshape(-1)
        # random garbage past bit n must be ignored
        pad = rng.integers(0, 256, 64, dtype=np.uint8)
        if n % 8:
            bv_h[-1] |= np.uint8((0xFF << (n % 8)) & 0xFF)
        exp = B.positions(np.concatenate([bv_h, pad]), n, base=11)
        for off in (0, 4):
            buf = torch.zeros(off + nbytes + 64 + 16, dtype=torch.uint8, device="cuda")
            buf[off:off + nbytes + 64] = torch.from_numpy(np.concatenate([bv_h, pad])).cuda()
            pos, num = sss.positions(buf[off:], n, base=11)
            k = int(num.item())
            assert k == len(exp), (n, density, off)
            assert (host(pos)[:k].astype(np.uint64) == exp).all(), (n, density, off)


@pytest.mark.parametrize("w", [14, 15])
@pytest.mark.parametrize("K", [2, 5, 8])
def test_one_code_byte_table_long_columns(sss, w, K):
    """14 and 15 bit with K <= 8 switch from the bucket table to the one-code byte table (narrow.cu) once
    the column has 2^24 codes; the short columns of the other tests never get there.  Breakpoints at 0-
    and top-adjacent codes, equality, empty and full predicates, ragged length, both layouts, counts."""
    n = (1 << 24) + 12_345
    rng = np.random.default_rng(77 * w + K)
    m = (1 << w) - 1
    codes = rng.integers(0, m + 1, size=n, dtype=np.uint32)
    codes[::7] = rng.integers(0, 9, size=len(codes[::7]), dtype=np.uint32)          # crowd the bottom of the domain
    codes[3::7] = m - rng.integers(0, 9, size=len(codes[3::7]), dtype=np.uint32)    # ... and the top
    packed = B.pack(codes, w)
    lo = rng.integers(0, m + 1, size=K).astype(np.uint64)
    hi = np.minimum(lo + rng.integers(0, (m + 1) // 4, size=K).astype(np.uint64), m)
    lo[0], hi[0] = 1, 1                                  # equality right above 0
    lo[1], hi[1] = 0, 3                                  # starts at 0: no breakpoint of its own below
    if K > 2:
        lo[2], hi[2] = m, m                              # the top code only
        lo[3], hi[3] = 0, m                              # everything
        lo[4], hi[4] = 9, 2                              # empty
    exp, ecnt = B.shared_scan(packed, n, w, lo, hi, granule=1)
    bv, cnt = sss.shared_scan(dev(packed), n, w, lo, hi)
    nb = (n + 7) // 8
    assert (host(bv)[:, :nb] == exp).all()
    assert (host(cnt).astype(np.uint64) == ecnt).all()
    lexp, lcnt = B.shared_scan_linear(packed, n, w, lo, hi, granule=1)
    lin, cnt2 = sss.shared_scan_linear(dev(packed), n, w, lo, hi)
    assert (host(lin)[: lexp.shape[0]] == lexp).all()
    assert (host(cnt2).astype(np.uint64) == lcnt).all()
