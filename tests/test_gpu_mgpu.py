"""The multi-GPU entry points of the C ABI (sss_mgpu_*, csrc/mgpu.cu) on however many GPUs the box
has (1 on the driver's test box, 2/4/8 under `gpurun --gpus N`): range partition, one ncclAllReduce
of the per-query counts, positions rebased to global code indices -- against the CPU oracle run on the
WHOLE column."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import binding as B  # noqa: E402


@pytest.fixture(scope="module")
def mg():
    from shared_simd_scan_b200 import mgpu
    assert torch.cuda.is_available()
    m = mgpu.MultiGpu(n_devices=min(torch.cuda.device_count(), 8))
    yield m
    m.close()


def host(t):
    return t.cpu().numpy()


def shards_of(mg, codes, w):
    n = len(codes)
    out = []
    for g, dev in enumerate(mg.devices):
        first, count = mg.shard(n, g)
        out.append(torch.from_numpy(B.pack(codes[first:first + count], w)).to(f"cuda:{dev}"))
    return out


@pytest.mark.parametrize("w,n", [(9, 1_000_003), (23, 300_007), (1, 4_099), (32, 70_001), (17, 100)])
def test_mgpu_scan_counts_bitvector_positions(mg, w, n):
    G = len(mg)
    codes = B.generate_codes(n, w, seed=1234 + w)
    packed_global = B.pack(codes, w)
    m = (1 << w) - 1
    d_shards = shards_of(mg, codes, w)
    for lo, hi in [(m // 5, m // 2), (int(codes[n // 3]), int(codes[n // 3])), (3, 2)]:
        exp, ecnt = B.scan_range(packed_global, n, w, lo, hi, granule=1)
        epos = B.positions(exp, n)
        bv, cnt, pos, num = mg.scan(d_shards, n, w, lo, hi, with_positions=True)
        pieces, allpos = [], []
        for g in range(G):
            first, count = mg.shard(n, g)
            assert first % 128 == 0 or count == 0
            assert int(cnt[g].item()) == ecnt, "every GPU holds the GLOBAL count after the all-reduce"
            pieces.append(host(bv[g])[: (count + 7) // 8])
            k = int(num[g].item())
            allpos.append(host(pos[g])[:k].astype(np.uint64))
        assert (np.concatenate(pieces)[: (n + 7) // 8] == exp).all()
        assert (np.concatenate(allpos) == epos).all(), "device order must be globally ascending"


@pytest.mark.parametrize("w,K,linear", [(23, 64, False), (17, 64, True), (9, 1024, True), (11, 8, False), (5, 1, False)])
def test_mgpu_shared_scan(mg, w, K, linear):
    G = len(mg)
    n = 200_003
    codes = B.generate_codes(n, w, seed=77 + w)
    packed_global = B.pack(codes, w)
    m = (1 << w) - 1
    rng = np.random.default_rng(w * 31 + K)
    lo = rng.integers(0, m + 1, size=K, dtype=np.uint64)
    hi = np.minimum(lo + rng.integers(0, max(2, 2 * (m + 1) // K), size=K, dtype=np.uint64), m)
    lo[0] = hi[0] = int(codes[5])
    d_shards = shards_of(mg, codes, w)
    outs, cnts = mg.shared_scan(d_shards, n, w, lo, hi, linear=linear)
    if linear:
        exp, ecnt = B.shared_scan_linear(packed_global, n, w, lo, hi, granule=1)
        rows = [host(outs[g])[: (mg.shard(n, g)[1] + 7) // 8] for g in range(G)]
        assert (np.concatenate(rows)[: (n + 7) // 8] == exp).all()
    else:
        exp, ecnt = B.shared_scan(packed_global, n, w, lo, hi, granule=1)
        got = np.concatenate([host(outs[g])[:, : (mg.shard(n, g)[1] + 7) // 8] for g in range(G)], axis=1)
        assert (got[:, : (n + 7) // 8] == exp).all()
    for g in range(G):
        assert (host(cnts[g]).astype(np.uint64) == ecnt).all(), "GLOBAL per-query counts on every GPU"


@pytest.mark.parametrize("tail", [1, 8, 32])
def test_mgpu_host_buffers(mg, tail):
    """One host column, every GPU pulling its own shard: equal to the single-GPU host entry point."""
    from shared_simd_scan_b200 import api
    w, n = 13, 3_000_017
    codes = B.generate_codes(n, w, seed=5)
    packed = B.pack(codes, w)
    lo, hi = 100, 4000
    out1, hits1 = api.host_scan(packed, n, w, lo, hi, tail_granule=tail)
    outN, hitsN = mg.host_scan(packed, n, w, lo, hi, tail_granule=tail)
    assert hits1 == hitsN and (out1 == outN).all()
    los, his = [lo, 0, 7, int(codes[9])], [hi, 8191, 3, int(codes[9])]
    s1, c1 = api.host_shared_scan(packed, n, w, los, his, tail_granule=tail)
    sN, cN = mg.host_shared_scan(packed, n, w, los, his, tail_granule=tail)
    assert (c1 == cN).all() and (s1 == sN).all()
