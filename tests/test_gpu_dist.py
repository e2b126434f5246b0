"""The one-process-per-GPU path (shared_simd_scan_b200.dist over torch.distributed) with the CUDA
kernels on real devices: two ranks, the count all-reduce, global counts against the CPU oracle run on
the WHOLE column, rank-order positions globally ascending.  With two or more GPUs the ranks use
different devices and NCCL; on a one-GPU box both ranks share cuda:0 and the 8-byte all-reduce goes
through gloo (NCCL refuses two ranks on one device) -- the kernels, the partitioning and the
rebasing are the same code."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SSS_ROOT"])
from oracle import binding as B
from shared_simd_scan_b200 import api, dist as sdist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
multi = torch.cuda.device_count() >= world
torch.cuda.set_device(rank if multi else 0)
dist.init_process_group("nccl" if multi else "gloo")
w, n = 23, 1_000_003
m = (1 << w) - 1
shard = sdist.generate_shard(n, w, seed=99)                      # this rank's range of the global column
codes = B.generate_codes(n, w, seed=99)                          # the whole column, for the oracle
packed = B.pack(codes, w)
lo, hi = m // 7, m // 3
bv, total = sdist.scan(shard, lo, hi)
exp, ecnt = B.scan_range(packed, n, w, lo, hi, granule=1)
assert int(total.item()) == ecnt, (int(total.item()), ecnt)       # GLOBAL count on every rank
first, count = shard.first, shard.count
mine = bv.cpu().numpy()[: (count + 7) // 8]
assert (mine[: count // 8] == exp[first // 8: first // 8 + count // 8]).all()
pos, num = sdist.positions(shard, bv)
allpos = sdist.gather_positions(pos, int(num.item()))
epos = B.positions(exp, n)
assert (allpos.cpu().numpy().astype(np.uint64) == epos).all()     # rank order == globally ascending
K = 64
rng = np.random.default_rng(7)
los = rng.integers(0, m + 1, size=K, dtype=np.uint64)
his = np.minimum(los + rng.integers(0, 2 * (m + 1) // K, size=K, dtype=np.uint64), m)
for linear in (False, True):
    out, counts = sdist.shared_scan(shard, los, his, linear=linear)
    _, ecounts = B.shared_scan(packed, n, w, los, his, granule=1)
    assert (counts.cpu().numpy().astype(np.uint64) == ecounts).all()
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank %d passed over %s\n" % (rank, "nccl" if multi else "gloo"))
sys.stdout.flush()
'''


def test_two_ranks_cuda_kernels_allreduce_positions(tmp_path):
    import torch
    assert torch.cuda.is_available()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SSS_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count("passed over") == 2, r.stdout
