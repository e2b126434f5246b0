"""Multi-GPU host logic (shared_simd_scan_b200.dist) at world_size 2 over gloo on CPU: range
partitioning, the per-query count all-reduce and position rebasing.  The per-shard compute is
injected (the CPU oracle stands in for the CUDA operators, which need a GPU); everything else is
the code that runs on the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as B
from shared_simd_scan_b200 import dist as sdist


def test_shard_range_properties():
    for n in (0, 1, 127, 128, 129, 1000, 10_000_019, 1 << 34):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                first, count = sdist.shard_range(n, world, r)
                assert first == covered and first % 128 == 0 or count == 0
                covered += count
            assert covered == n
    with pytest.raises(ValueError):
        sdist.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_ops(codes_global, w):
    """CPU stand-ins with the signatures of api.scan / api.shared_scan / api.positions."""
    def scan_op(packed, n, w_, lo, hi):
        bv, cnt = B.scan_range(packed, n, w_, lo, hi, granule=1)
        return torch.from_numpy(bv), torch.tensor([cnt], dtype=torch.int64)

    def shared_op(packed, n, w_, lo, hi):
        outs, counts = B.shared_scan(packed, n, w_, lo, hi, granule=1)
        return torch.from_numpy(outs), torch.from_numpy(counts.astype(np.int64))

    def positions_op(bitvec, n, base=0, capacity=None):
        pos = B.positions(bitvec.numpy(), n, base=base)
        return torch.from_numpy(pos.astype(np.int64)), len(pos)

    return scan_op, shared_op, positions_op


def _worker(rank, world, port, w, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        codes = B.generate_codes(n, w, seed=11)
        first, count = sdist.shard_range(n, world, rank)
        shard = sdist.Shard(B.pack(codes[first:first + count], w), n, first, count, w)
        scan_op, shared_op, positions_op = _oracle_ops(codes, w)
        m = (1 << w) - 1
        lo, hi = m // 5, m // 2

        bv, total = sdist.scan(shard, lo, hi, scan_op=scan_op)
        truth = (codes >= lo) & (codes <= hi)
        assert int(total.item()) == int(truth.sum())           # GLOBAL count on every rank
        local_bits = np.unpackbits(bv.numpy(), bitorder="little")[:count]
        assert (local_bits == truth[first:first + count]).all()  # this rank's slice of the bitvector

        pos, num = sdist.positions(shard, bv, positions_op=positions_op)
        allpos = sdist.gather_positions(pos, num)
        assert (allpos.numpy() == np.nonzero(truth)[0]).all()   # rank order == globally ascending

        los, his = [lo, 0, 7, int(codes[5])], [hi, m, 3, int(codes[5])]
        _, counts = sdist.shared_scan(shard, los, his, shared_op=shared_op)
        exp = [int(((codes >= a) & (codes <= b)).sum()) for a, b in zip(los, his)]
        assert counts.tolist() == exp
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("w,n", [(9, 10_007), (17, 300), (1, 129)])
def test_world2_gloo(w, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
