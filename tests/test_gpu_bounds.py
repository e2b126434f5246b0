"""Write bounds of the device entry points.  compute-sanitizer is closed on the GPU pool, so memory
safety of the stores is checked here the blunt way: every output buffer is a view of EXACTLY the size
the C ABI documents (sss_bitvec_bytes, 4 bytes per code, sss_packed_bytes, capacity x 8, K x stride,
rows x K) inside a larger allocation filled with a canary byte, ragged sizes at every width class, and
the canaries before and after the view must survive the call.  (Results are compared with the oracle
in test_gpu_parity.py; this file only looks at what must NOT be written.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import binding as B  # noqa: E402

CANARY = 0xA5
PAD = 4096
WIDTHS = [1, 2, 3, 5, 8, 9, 11, 12, 13, 14, 16, 17, 23, 24, 28, 31, 32]
SIZES = [1, 33, 1000, 4096 * 32 + 5, 300_007]


@pytest.fixture(scope="module")
def sss():
    import shared_simd_scan_b200 as m
    from shared_simd_scan_b200 import api
    assert torch.cuda.is_available()
    m.Context.current()
    return api


class Guarded:
    """`nbytes` usable bytes (256-byte aligned start) between two canary zones."""

    def __init__(self, nbytes, dtype=torch.uint8):
        self.nbytes = int(nbytes)
        self.buf = torch.full((PAD + self.nbytes + PAD,), CANARY, dtype=torch.uint8, device="cuda")
        self.view = self.buf[PAD:PAD + self.nbytes]
        if dtype != torch.uint8:
            self.view = self.view.view(dtype)

    def check(self, what, written_upto=None):
        torch.cuda.synchronize()
        assert bool((self.buf[:PAD] == CANARY).all()), f"{what}: wrote BEFORE the buffer"
        assert bool((self.buf[PAD + self.nbytes:] == CANARY).all()), f"{what}: wrote PAST the buffer"
        if written_upto is not None:  # bytes of the view the call has no business touching
            assert bool((self.buf[PAD + int(written_upto):PAD + self.nbytes] == CANARY).all()), \
                f"{what}: wrote past its documented extent inside the buffer"


def column(n, w, seed):
    codes = B.generate_codes(n, w, seed=seed)
    return codes, torch.from_numpy(B.pack(codes, w)).cuda()


@pytest.mark.parametrize("w", WIDTHS)
def test_unpack_scan_pack_write_bounds(sss, w):
    m = (1 << w) - 1
    for n in SIZES:
        codes, packed = column(n, w, seed=900 + w)
        g = Guarded(4 * n, torch.int32)  # exactly one int32 per code
        sss.decompress(packed, n, w, out=g.view)
        g.check(f"unpack w={w} n={n}")
        g = Guarded(sss.bitvec_bytes(n))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        sss.scan(packed, n, w, m // 4, m // 2, out=g.view, count=cnt)
        g.check(f"scan w={w} n={n}", written_upto=4 * ((n + 31) // 32))
        g = Guarded(sss.packed_bytes(n, w))
        sss.compress(torch.from_numpy(codes.view(np.int32)).cuda(), w, out=g.view)
        g.check(f"pack w={w} n={n}")


@pytest.mark.parametrize("K", [1, 3, 8, 64, 200, 1024])
@pytest.mark.parametrize("w", [1, 5, 9, 12, 13, 14, 17, 24, 29, 32])
def test_shared_scan_write_bounds(sss, w, K):
    m = (1 << w) - 1
    rng = np.random.default_rng(w * 1000 + K)
    lo = rng.integers(0, m + 1, K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, max(2, (m + 1) // 4), K, dtype=np.uint64), m)
    for n in [33, 1000, 4096 * 32 + 5, 70_001]:
        _, packed = column(n, w, seed=950 + w)
        words = 4 * ((n + 31) // 32)
        stride = (words + 127) // 128 * 128
        g = Guarded(K * stride)
        out = g.view.view(K, stride)
        sss.shared_scan(packed, n, w, lo.tolist(), hi.tolist(), out=out)
        g.check(f"shared w={w} K={K} n={n}")
        torch.cuda.synchronize()
        assert bool((out[:, words:] == CANARY).all()), f"shared w={w} K={K} n={n}: wrote past a bitvector's words"
        rows = 4 * ((n + 31) // 32)
        g = Guarded(rows * K)
        sss.shared_scan_linear(packed, n, w, lo.tolist(), hi.tolist(), out=g.view.view(rows, K))
        g.check(f"linear w={w} K={K} n={n}")


@pytest.mark.parametrize("density", [0.0005, 0.02, 0.3, 0.7, 1.0])
def test_positions_write_bounds(sss, density):
    rng = np.random.default_rng(int(density * 1e4))
    for n in [77, 32768 * 3 + 11, (1 << 20) + 5]:
        bits = rng.random(((n + 7) // 8) * 8) < density
        bv = torch.from_numpy(np.packbits(bits.reshape(-1, 8)[:, ::-1], axis=1).reshape(-1)).cuda()
        hits = int(bits[:n].sum())
        for cap in {hits, hits // 2, 1}:
            g = Guarded(max(cap, 1) * 8, torch.int64)
            pos, num = sss.positions(bv, n, capacity=cap, out=g.view)
            assert int(num.item()) == hits
            g.check(f"positions density={density} n={n} capacity={cap}", written_upto=min(cap, hits) * 8)
