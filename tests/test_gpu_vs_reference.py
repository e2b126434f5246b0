"""The CUDA path against the UNMODIFIED reference itself (oracle/_ref/libref_wNN.so, built from
/root/reference by oracle/build_ref.sh and shipped to the GPU box) at BASELINE.json's configurations,
scaled where the CPU would take too long.  Complements the oracle-based parity tests: here nothing of
ours sits between the two results except the packed input bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import binding as B  # noqa: E402

needs_ref = pytest.mark.skipif(not B.ref_available(11), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def sss():
    from shared_simd_scan_b200 import api
    api.Context.current()
    return api


def random_column(n, w, seed):
    rng = np.random.default_rng(seed)
    packed = np.zeros(B.compressed_buffer_size(w, n), dtype=np.uint8)
    packed[: n * w // 8] = np.frombuffer(rng.bytes(n * w // 8), dtype=np.uint8)  # uniform bits = uniform codes
    return packed


@needs_ref
def test_config0_single_scan_16m_11bit(sss):
    # BASELINE configs[0]: single predicate over 16M 11-bit codes, the reference AVX2 scan
    w, n = 11, 1 << 24
    packed = random_column(w=w, n=n, seed=1)
    key = int(B.unpack(packed, 8, w)[3])
    ref, hits = B.ref_scan("256_unrolled", key, packed, n, w)
    out, got_hits = sss.host_scan(packed, n, w, key, tail_granule=32)
    assert (out == ref).all() and got_hits == hits
    bv, cnt = sss.scan(torch.from_numpy(packed).cuda(), n, w, key)
    assert (bv.cpu().numpy()[: n // 8] == ref[: n // 8]).all() and int(cnt.item()) == hits


@needs_ref
@pytest.mark.parametrize("w", [1, 4, 9, 13, 16, 17, 23, 26, 28, 32])
def test_unpack_and_scan_vs_reference(sss, w):
    # BASELINE configs[1] (decompression sweep) and the scan at the widths the SSE family is exact for
    n = (1 << 20) + 37
    packed = random_column(w=w, n=n, seed=w)
    assert (sss.host_decompress(packed, n, w) == B.ref_decompress("128_unrolled", packed, n, w)).all()
    key = int(B.unpack(packed, 1, w)[0])
    skey = key - (1 << 32) if key >= (1 << 31) else key
    ref, hits = B.ref_scan("128", skey, packed, n, w)
    out, got = sss.host_scan(packed, n, w, key, tail_granule=8)
    assert (out == ref).all() and got == hits


@needs_ref
@pytest.mark.parametrize("w,K,n", [(17, 64, 1 << 18), (23, 64, 1 << 18), (9, 8, 1 << 20)])
def test_shared_scan_vs_reference(sss, w, K, n):
    # BASELINE configs[2] / [4] scaled: K equality predicates in one pass, reference shared_scan_128_standard
    packed = random_column(w=w, n=n, seed=100 + w)
    codes = B.unpack(packed, 4096, w).astype(np.uint32)
    keys = [int(c) for c in codes[:K]]
    skeys = [k - (1 << 32) if k >= (1 << 31) else k for k in keys]
    ref = B.ref_shared_scan("128_standard", skeys, packed, n, w)
    outs, counts = sss.host_shared_scan(packed, n, w, keys, tail_granule=8)
    assert (outs == ref).all()
    assert [int(c) for c in counts] == [int(np.unpackbits(ref[k]).sum()) for k in range(K)]


@needs_ref
def test_config3_linear_1024_vs_reference(sss):
    # BASELINE configs[3] scaled: 1024 predicates over 9-bit codes, linear layout, reference linear_simple
    w, K, n = 9, 1024, 1 << 15
    packed = random_column(w=w, n=n, seed=3)
    keys = [(i * 5) % 512 for i in range(K)]
    ref = B.ref_shared_scan_linear("linear_simple", keys, packed, n, w)
    lin, _ = sss.host_shared_scan_linear(packed, n, w, keys, tail_granule=8)
    assert (lin == ref).all()
