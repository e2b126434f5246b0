#!/usr/bin/env python
"""bench.py -- the measurement contract for shared_simd_scan_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload ("step"): one pass of the hot path over one batch of synthetic input =
the fused unpack + range-predicate scan (bitvector + match count out) over a uniform-random column
of 2^30 codes at EVERY bit width 1..32 -- the configuration BASELINE.json's target is quoted on
("single-GPU, 1B codes, fused unpack+scan at every bit width 1-32").  32 kernel launches per step.
With N GPUs every rank scans its own range shard of a column of N * 2^30 codes per width (weak
scaling), and the per-width match counts are combined with one small NCCL allreduce per step.

value   = codes scanned by all ranks / max-over-ranks device time (CUDA events, inputs resident in HBM)
e2e     = the same sweep through the host-buffer C-ABI call the C++ drop-in layer uses
          (sss_host_scan_range: pinned host buffers, H2D + kernel + D2H inside the timed region)
roofline= algorithmic bytes (w/8 + 1/8 per code) / measured kernel time vs. MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference = the reference's own AVX2 scan (oracle/_ref) on the host cores.
The line stays compact (< 6 KB, same keys at every N): `summary` carries a handful of per-K numbers
(shared scan K = 64 at 17 / 23 bit, linear K = 1024 at 9 bit, cfg 5's sharded shared scan with its count
all-reduce, ...), `per_width_frac` the 32 per-width roofline fractions; the full tables (per_width,
shared_sweep, unpack, pack, positions, calibration, cfg1 CUDA-graph timing, cfg4 at 2^30 codes, the CPU
reference detail) go to the side file named by `extras_file` (gpurun_out/bench_extras_N<gpus>.json).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

_json_out = sys.stdout


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the process's original stdout (see main())."""
    _json_out.write(json.dumps(line) + "\n")
    _json_out.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTHS = list(range(1, 33))
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def captured_traffic():
    """DRAM bytes of the 32 scan launches of one step, from the committed ncu capture
    (profiles/r02_scan_traffic_v2.json; tools/scan_traffic.py + tools/summarise_traffic.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_scan_traffic_v2.json")) as f:
            t = json.load(f)
        return float(t["traffic_per_step"]), float(t["algorithmic_per_step"])
    except Exception:
        return None, None


def predicate(w):
    """Range predicate of the sweep: [m/4, m/2] (about 25 % selectivity for w >= 3)."""
    m = (1 << w) - 1
    return m // 4, m // 2


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if (t0 is None or t >= t0 - 0.05) and (t1 is None or t <= t1 + 0.15)]
        if not rows:
            rows = [l for (_, l) in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except Exception:
                continue
            for name, val in zip(names, p[2:]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY places that execute anything under oracle/)
# ---------------------------------------------------------------------------------------------

_cpu_columns = {}


def _cpu_column(w, logn, seed=42):
    """Uniform-random packed column on the host (uniform bits == uniform codes), cached per (w, logn)
    so that the steps of one run scan the same data instead of regenerating 4 GiB of it per step."""
    import numpy as np
    from oracle import binding as B
    key = (w, logn)
    if key not in _cpu_columns:
        n = 1 << logn
        rng = np.random.default_rng(seed + w)
        packed = np.zeros(B.compressed_buffer_size(w, n), dtype=np.uint8)
        payload = n * w // 8
        # numpy draws ~1 GB/s: columns beyond 64 MiB repeat a 64 MiB random block (the scan's speed does
        # not depend on the values; the sample only has to be too large for the caches)
        block = np.frombuffer(rng.bytes(min(payload, 64 << 20)), dtype=np.uint8)
        for off in range(0, payload, len(block)):
            m = min(len(block), payload - off)
            packed[off:off + m] = block[:m]
        _cpu_columns[key] = packed
    return _cpu_columns[key]


def cpu_reference_sweep(logn, reps, threads):
    """The reference's own CPU scan over a bounded sample: 2^logn uniform codes per width.

    oracle/_ref (the unmodified reference compiled per width) when present: scan_256_unrolled for
    w <= 16 (the "reference AVX2 scan"), scan_128_unrolled above (the AVX path is only exact up to
    16 bits); every host thread scans its own range of the column.  If
    oracle/_ref is absent the plain-C oracle port is timed instead (kind "port").
    Returns (codes_per_s over the whole sweep, per-width seconds, kind)."""
    import ctypes as C

    from oracle import binding as B
    n = 1 << logn
    per_w = {}
    kind = "reference" if all(B.ref_available(w) for w in WIDTHS) else "port"
    for w in WIDTHS:
        packed = _cpu_column(w, logn)
        key = 3 & ((1 << w) - 1)  # the reference only has equality predicates (src/simd_scan.cpp:47)
        if kind == "reference":
            variant = B.SCAN["256_unrolled"] if w <= 16 else B.SCAN["128_unrolled"]
            hits = C.c_longlong(0)
            s = B.ref(w).ref_time_scan(variant, key, packed.ctypes.data, n, threads, reps, C.byref(hits))
        else:
            s, _ = B.time_scan_range(packed, n, w, key, key, threads, reps)
        per_w[w] = s
    total = sum(per_w.values())
    return len(WIDTHS) * n / total, per_w, kind


def cpu_reference_shared(threads, w=17, K=64, logn=22, reps=2):
    """The reference's shared scans (K equality predicates, one pass) on the host cores, column
    range-partitioned over the threads -- the CPU side of `e2e_shared`.  None when oracle/_ref is absent."""
    import numpy as np
    from oracle import binding as B
    if not B.ref_available(w):
        return None
    n = 1 << logn
    packed = _cpu_column(w, logn, seed=7)
    keys = np.ascontiguousarray(np.random.default_rng(7).integers(0, 1 << w, size=K), dtype=np.int32)
    sec_std = B.ref(w).ref_time_shared_scan(B.SHARED["128_standard"], keys.ctypes.data, K, packed.ctypes.data, n,
                                            threads, reps)
    sec_lin = B.ref(w).ref_time_shared_scan_linear(B.LINEAR["linear_standard"], keys.ctypes.data, K,
                                                   packed.ctypes.data, n, threads, reps)
    return {"w": w, "K": K, "log2_n": logn, "threads": threads,
            "shared_scan_128_standard_codes_per_s": n / sec_std,
            "shared_scan_128_linear_standard_codes_per_s": n / sec_lin}


def cpu_reference_detail(threads):
    """The other BASELINE.json configs on the host cores, scaled: codes/s is
    size-independent beyond cache.  Unmodified reference via oracle/_ref; [] when it is not built."""
    import ctypes as C

    import numpy as np
    from oracle import binding as B
    if not all(B.ref_available(w) for w in (9, 11, 17)):
        return []
    rng = np.random.default_rng(7)
    rows = []

    def column(w, n):
        packed = np.zeros(B.compressed_buffer_size(w, n), dtype=np.uint8)
        packed[: n * w // 8] = np.frombuffer(rng.bytes(n * w // 8), dtype=np.uint8)
        return packed

    # cfg 1: single predicate, 2^24 x 11 bit, the reference AVX2 scan, one thread and all threads
    w, n = 11, 1 << 24
    packed = column(w, n)
    hits = C.c_longlong(0)
    for t in (1, threads):
        sec = B.ref(w).ref_time_scan(B.SCAN["256_unrolled"], 3, packed.ctypes.data, n, t, 5, C.byref(hits))
        rows.append({"config": "cfg1 scan 2^24 x 11 bit", "variant": "scan_256_unrolled", "threads": t,
                     "gcodes_per_s": round(n / sec / 1e9, 2)})
    # cfg 2 sample: decompression
    for w, variant in ((9, "256_avx2"), (17, "128_unrolled")):
        n = 1 << 25
        packed = column(w, n)
        sec = B.ref(w).ref_time_decompress(B.DECOMPRESS[variant], packed.ctypes.data, n, threads, 3)
        rows.append({"config": f"cfg2 unpack 2^25 x {w} bit", "variant": "decompress_" + variant, "threads": threads,
                     "gcodes_per_s": round(n / sec / 1e9, 2)})
    # cfg 3 / 5 scaled: shared scan K = 64
    w, n, K = 17, 1 << 22, 64
    packed = column(w, n)
    keys = np.ascontiguousarray(rng.integers(0, 1 << w, size=K), dtype=np.int32)
    for name, fn in (("shared_scan_128_standard", lambda: B.ref(w).ref_time_shared_scan(
            B.SHARED["128_standard"], keys.ctypes.data, K, packed.ctypes.data, n, threads, 2)),
                     ("shared_scan_128_linear_standard", lambda: B.ref(w).ref_time_shared_scan_linear(
                         B.LINEAR["linear_standard"], keys.ctypes.data, K, packed.ctypes.data, n, threads, 2))):
        rows.append({"config": "cfg3 shared scan K=64, 2^22 x 17 bit", "variant": name, "threads": threads,
                     "gcodes_per_s": round(n / fn() / 1e9, 3)})
    # cfg 4 scaled: linear shared scan K = 1024, 9 bit
    w, n, K = 9, 1 << 19, 1024
    packed = column(w, n)
    keys = np.ascontiguousarray(np.arange(K) % 512, dtype=np.int32)
    sec = B.ref(w).ref_time_shared_scan_linear(B.LINEAR["linear_standard"], keys.ctypes.data, K, packed.ctypes.data, n,
                                               threads, 2)
    rows.append({"config": "cfg4 linear shared scan K=1024, 2^19 x 9 bit", "variant": "shared_scan_128_linear_standard",
                 "threads": threads, "gcodes_per_s": round(n / sec / 1e9, 4)})
    return rows


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    logn = args.ref_logn
    # warm-up + timed steps: each step = the sweep over the bounded sample (columns generated once)
    for _ in range(args.warmup):
        cpu_reference_sweep(logn, 1, threads)
    t_steps = []
    kind = "port"
    for _ in range(args.steps):
        v, per_w, kind = cpu_reference_sweep(logn, 1, threads)
        t_steps.append(sum(per_w.values()))
    sec = statistics.mean(t_steps)
    value = len(WIDTHS) * (1 << logn) / sec
    sample = (f"bounded sample: 2^{logn} uniform codes per width (workload: 2^{args.logn}), equality key 3; "
              f"{'scan_256_unrolled (w<=16) / scan_128_unrolled (w>16) of the unmodified reference' if kind == 'reference' else 'plain-C oracle port'}"
              f", column range-partitioned over {threads} host threads")
    line = {
        "impl": "reference", "metric": "scanned_codes_per_s", "value": value, "unit": "codes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "codes/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "codes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if kind == "reference" and not args.no_extras:
        # the CPU side of the product arm's `e2e_shared` (shared scan, K = 64 equality predicates, 17 bit)
        sh = cpu_reference_shared(threads)
        if sh:
            line["e2e_shared"] = {"value": max(sh["shared_scan_128_standard_codes_per_s"],
                                               sh["shared_scan_128_linear_standard_codes_per_s"]),
                                  "unit": "codes/s", "workload": "shared scan K=64, 17 bit (best of the reference's "
                                  "shared_scan_128_standard / _linear_standard)", **{k: (round(v, 1) if isinstance(v, float) else v)
                                                                                   for k, v in sh.items()}}
    emit(line)
    return 0


def workload_config(args):
    return {"workload": f"fused unpack+range-scan sweep, bit widths 1-32, 2^{args.logn} uniform-random codes per "
                        "width per GPU, range predicate [m/4, m/2], bitvector + match count out",
            "codes_per_width_per_gpu": 1 << args.logn, "widths": "1-32", "predicate": "lo=m/4, hi=m/2 (m=2^w-1)",
            "l2": "inputs larger than L2; every step streams ~70 GB through 32 different columns",
            "parallelism": f"range-sharded column, {args.gpus} GPU(s), NCCL allreduce of match counts"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

class Dist:
    """The multi-rank plumbing of one bench process (one process per GPU, torch.distributed/NCCL)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: shared_simd_scan_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        else:
            self.dist = None

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def all_reduce(self, t):
        if self.dist:
            self.dist.all_reduce(t)
        return t

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def random_ranges(torch, w, K, seed):
    """K uniform-random ranges: lo uniform, width uniform in [0, 2 * 2^w / K)."""
    m = (1 << w) - 1
    g = torch.Generator().manual_seed(seed)
    lo = torch.randint(0, m + 1, (K,), generator=g)
    hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
    return lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()


def measure_cfg5(D, api, logn, steps, peak):
    """BASELINE configs[4]: shared scan, K = 64 range predicates over 23-bit codes, the global column
    range-sharded over the ranks (2^logn codes per GPU; 2^31 x 8 GPUs = 16 Gi codes), per-query counts
    combined with ONE NCCL all-reduce inside the timed region.  Checks the all-reduced counts against
    the per-rank counts and every rank's count 0 against the popcount of its own bitvector 0."""
    torch = D.torch
    from shared_simd_scan_b200 import dist as sdist
    w, K = 23, 64
    shard = sdist.generate_shard((1 << logn) * D.world, w, seed=42 + w)
    lo_d, hi_d = random_ranges(torch, w, K, seed=42)
    out = torch.empty((K, (api.bitvec_bytes(shard.count) + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
    counts = torch.empty(K, dtype=torch.int64, device="cuda")

    def step():
        api.shared_scan(shard.packed, shard.count, w, lo_d, hi_d, out=out, counts=counts)
        D.all_reduce(counts)

    # correctness of the collective: global == sum of locals; local count 0 == popcount of bitvector 0
    api.shared_scan(shard.packed, shard.count, w, lo_d, hi_d, out=out, counts=counts)
    local = counts.clone()
    words = out[0, : shard.count // 8].view(torch.int64)
    pop0 = 0
    for part in words.split(1 << 24):  # SWAR popcount, chunked to bound the temporaries
        x = part.clone()
        x = x - ((x >> 1) & 0x5555555555555555)
        x = (x & 0x3333333333333333) + ((x >> 2) & 0x3333333333333333)
        x = (x + (x >> 4)) & 0x0F0F0F0F0F0F0F0F
        pop0 += int(((x * 0x0101010101010101) >> 56 & 0xFF).sum().item())
    assert pop0 == int(local[0].item()), "cfg5: count 0 != popcount of bitvector 0"
    D.all_reduce(counts)
    if D.dist:
        gathered = [torch.empty_like(local) for _ in range(D.world)]
        D.dist.all_gather(gathered, local)
        assert torch.equal(torch.stack(gathered).sum(0), counts), "cfg5: all-reduced counts != sum of per-rank counts"
    for _ in range(3):
        step()
    D.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    D.barrier()
    ms = D.max_over_ranks(a.elapsed_time(b)) / steps
    bytes_per_gpu = shard.count * (w / 8 + K / 8)
    res = {"w": w, "K": K, "log2_n_per_gpu": logn, "ms": round(ms, 4),
           "gcodes_per_s": round(D.world * shard.count / ms / 1e6, 1),
           "gb_per_s_per_gpu": round(bytes_per_gpu / ms / 1e6, 1),
           "frac": round(bytes_per_gpu / ms / 1e6 / peak, 4), "allreduce_in_timed_region": D.world > 1,
           "count0": int(counts[0].item())}
    del out, shard
    return res


def measure_e2e_scan(D, api, columns, n, e2e_steps):
    """The contract's e2e: the 32-width sweep through sss_host_scan_range from pinned host memory
    (H2D copy of every column + kernel + D2H of bitvector and count inside the timed region)."""
    torch = D.torch
    h_in = api.pinned_empty(api.packed_bytes(n, 32))
    h_out = api.pinned_empty(api.bitvec_bytes(n))
    e2e_sec, h2d, d2h, counts = [], 0, 0, []
    for s in range(e2e_steps + 1):  # first pass = warm-up (staging buffers get allocated)
        sec = 0.0
        h2d = d2h = 0
        for w in WIDTHS:
            nb = api.packed_bytes(n, w)
            h_in[:nb].copy_(columns[w][:nb])  # refill the host column (untimed)
            D.barrier()  # every rank has refilled: no rank's untimed D2H competes with another's timed H2D
            lo, hi = predicate(w)
            t0 = time.perf_counter()
            _, c = api.host_scan(h_in, n, w, lo, hi, tail_granule=1, out=h_out)
            sec += time.perf_counter() - t0
            D.barrier()
            h2d += (n // 32) * w * 4
            d2h += n // 8 + 8
            if s == e2e_steps:
                counts.append(c)
        if s > 0:
            e2e_sec.append(sec)
    sec = D.max_over_ranks(statistics.mean(e2e_sec))
    del h_in, h_out
    return {"value": D.world * len(WIDTHS) * n / sec, "unit": "codes/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": e2e_steps, "h2d_gb_per_s_per_gpu": round(h2d / sec / 1e9, 1),
            "api": "sss_host_scan_range (pinned host buffers, 64 MiB chunks, 3 streams)"}, counts


def measure_e2e_shared(D, api, columns, n, steps, w=17, K=64):
    """e2e of BASELINE configs[2]: shared scan, K = 64 range predicates over 2^30 17-bit codes through
    sss_host_shared_scan (host column in, K host bitvectors + counts out, all copies timed)."""
    import numpy as np
    torch = D.torch
    lo_d, hi_d = random_ranges(torch, w, K, seed=42 + K)
    lo_h = lo_d.cpu().numpy().view(np.uint32).copy()
    hi_h = hi_d.cpu().numpy().view(np.uint32).copy()
    nb = api.packed_bytes(n, w)
    h_in = api.pinned_empty(nb)
    h_in.copy_(columns[w][:nb])
    stride = (api.bitvec_bytes(n) + 127) // 128 * 128
    h_out = api.pinned_empty(K * stride)
    h_cnt = np.zeros(K, dtype=np.uint64)
    torch.cuda.synchronize()
    secs = []
    for s in range(steps + 1):
        t0 = time.perf_counter()
        api.host_shared_scan_into(h_in, n, w, lo_h, hi_h, h_out, stride, h_cnt)
        if s > 0:
            secs.append(time.perf_counter() - t0)
    sec = D.max_over_ranks(statistics.mean(secs))
    # same counts as the device-resident call
    _, cnt = api.shared_scan(columns[w], n, w, lo_d, hi_d)
    assert (cnt.cpu().numpy().astype(np.uint64) == h_cnt).all(), "e2e_shared: host path != device path counts"
    del h_in, h_out
    return {"value": D.world * n / sec, "unit": "codes/s", "workload": f"shared scan K={K}, {w} bit, 2^{int(math.log2(n))} codes per GPU",
            "h2d_bytes_per_step": (n // 32) * w * 4, "d2h_bytes_per_step": K * (n // 8) + 8 * K, "steps": steps,
            "d2h_gb_per_s_per_gpu": round(K * (n // 8) / sec / 1e9, 1), "api": "sss_host_shared_scan (pinned)"}


def run_ours(args):
    import numpy as np  # noqa: F401
    import torch

    from shared_simd_scan_b200 import api

    D = Dist()
    rank, world, local = D.rank, D.world, D.local
    n = 1 << args.logn
    peak, peak_src = measured_peak()

    # ---- resident inputs: one packed column per width, this rank's shard of the global column ----
    columns = {w: api.generate_packed(n, w, seed=42 + w, first_index=rank * n) for w in WIDTHS}
    bitvec = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    counts = torch.zeros(len(WIDTHS), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def step():
        for i, w in enumerate(WIDTHS):
            lo, hi = predicate(w)
            api.scan(columns[w], n, w, lo, hi, out=bitvec, count=counts[i:i + 1])
        D.all_reduce(counts)  # the only collective on the path: 32 x int64

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    D.barrier()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    wall0 = time.perf_counter()
    t_begin.record()
    for k in range(args.steps):
        step()
    t_end.record()
    D.barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop(wall0, wall1)
    ms_per_step = D.max_over_ranks(t_begin.elapsed_time(t_end)) / args.steps
    value = world * len(WIDTHS) * n / (ms_per_step * 1e-3)

    # roofline of the step's kernel (32 launches of scan_range_kernel<W>): algorithmic bytes of a step over the
    # device time of a step, both from the timed region above (so launch gaps count against the kernel)
    total_bytes = sum(n * (w / 8 + 1 / 8) for w in WIDTHS)
    achieved = total_bytes / (ms_per_step * 1e-3) / 1e9

    # the collective is right: the all-reduced counts are the sum of the ranks' own counts
    local_counts = torch.zeros(len(WIDTHS), dtype=torch.int64, device="cuda")
    for i, w in enumerate(WIDTHS):
        lo, hi = predicate(w)
        api.scan(columns[w], n, w, lo, hi, out=bitvec, count=local_counts[i:i + 1])
    final_counts = local_counts.cpu().tolist()
    if D.dist:
        gathered = [torch.empty_like(local_counts) for _ in range(world)]
        D.dist.all_gather(gathered, local_counts)
        assert torch.equal(torch.stack(gathered).sum(0), counts), "all-reduced counts != sum of per-rank counts"
    else:
        assert torch.equal(local_counts, counts)

    # per-width launch duration, after the timed region: `reps` back-to-back launches of one width between
    # two events on the launching stream (an event pair around a single 50 us launch would add ~20 %)
    per_width = []
    reps = 10
    for i, w in enumerate(WIDTHS):
        lo, hi = predicate(w)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        api.scan(columns[w], n, w, lo, hi, out=bitvec, count=counts[i:i + 1])
        a.record()
        for _ in range(reps):
            api.scan(columns[w], n, w, lo, hi, out=bitvec, count=counts[i:i + 1])
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / reps
        nb = n * (w / 8 + 1 / 8)
        per_width.append({"w": w, "ms": round(ms, 4), "gcodes_per_s": round(n / ms / 1e6, 1),
                          "gb_per_s": round(nb / ms / 1e6, 1), "frac_of_measured_peak": round(nb / ms / 1e6 / peak, 4),
                          "frac_of_8TBps": round(nb / ms / 1e6 / 8000.0, 4)})

    # ---- e2e: the host-buffer C-ABI calls (what include/simd_scan.hpp wraps), pinned host memory ----
    e2e_steps = max(1, args.e2e_steps)
    # the host side of a rank lives next to its GPU while it copies: pinned buffers are first-touched on the
    # GPU's NUMA node and no copy crosses the socket interconnect (matters from 4 ranks per box on); the
    # CPU baseline further down gets every core back
    affinity0 = os.sched_getaffinity(0)
    host_cpus = api.bind_host_thread(local)
    e2e, e2e_counts = measure_e2e_scan(D, api, columns, n, e2e_steps)
    e2e["host_cpus_per_rank"] = host_cpus
    if world == 1:
        assert e2e_counts == final_counts, "host-buffer path and device-resident path disagree on the match counts"
    e2e_shared = measure_e2e_shared(D, api, columns, n, e2e_steps) if not args.no_extras else None
    os.sched_setaffinity(0, affinity0)

    # ---- the compact per-K summary, at every N (per-GPU device-resident numbers: max time over ranks) ----
    summary = {}
    extras_out = {"per_width": per_width}
    if not args.no_extras:
        summary["cfg5_shared_k64_w23"] = measure_cfg5(D, api, min(args.cfg5_logn, args.logn + 1), 5, peak)
        for kind, w, K in (("shared", 17, 64), ("linear", 9, 64), ("shared", 17, 8)):
            summary[f"{kind}_k{K}_w{w}"] = shared_point(D, api, columns, n, kind, w, K, peak)
    fr = [p["frac_of_measured_peak"] for p in per_width]
    summary["scan_frac_min"] = {"w": WIDTHS[fr.index(min(fr))], "frac": min(fr)}
    summary["scan_widths_below_0.95"] = [p["w"] for p in per_width if p["frac_of_measured_peak"] < 0.95]

    traffic, _ = captured_traffic()
    line = {
        "metric": "scanned_codes_per_s", "value": value, "unit": "codes/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "e2e_shared": e2e_shared,
        "gpu_launches": args.steps * len(WIDTHS),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic if args.logn == 30 else None,
                     "algorithmic_bytes": total_bytes, "peak_source": peak_src,
                     "traffic_source": "ncu dram bytes of the step's 32 launches, profiles/r02_scan_traffic_v2.json",
                     "kernel": "sss::scan_range_kernel<W>, the 32 launches of a step (W = 1: scan_direct_kernel); "
                               "bytes = sum n*(w/8+1/8), time = the step's device time in the timed region",
                     "frac_of_8TBps_nominal": round(achieved / 8000.0, 4)},
        "per_width_frac": [round(f, 3) for f in fr],
        "summary": summary,
    }

    # ---- the full tables: measured on rank 0 of a single-GPU run, written to the side file ----
    if rank == 0 and world == 1 and not args.no_extras:
        extras_out.update(extras(args, api, columns, n, peak))
    threads = os.cpu_count() or 1
    if rank == 0 and not args.no_cpu:
        # bounded CPU sample; smaller when other ranks are waiting (the full sample runs at N = 1)
        cpu_logn = args.cpu_logn if world == 1 else min(args.cpu_logn, 24)
        t0 = time.perf_counter()
        v, per_w, kind = cpu_reference_sweep(cpu_logn, 2, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": "codes/s", "cores": threads, "kind": kind,
            "sample": f"same sweep on 2^{cpu_logn} uniform codes per width (best of 2), equality key 3; "
                      "scan_256_unrolled (w<=16) / scan_128_unrolled (w>16) of the unmodified reference, column "
                      f"range-partitioned over {threads} host threads; {time.perf_counter() - t0:.1f} s of wall time"}
        if world == 1 and not args.no_extras:
            extras_out["cpu_reference_detail"] = cpu_reference_detail(threads)
            sh = cpu_reference_shared(threads)
            if sh and e2e_shared:
                e2e_shared["cpu_reference_codes_per_s"] = round(
                    max(sh["shared_scan_128_standard_codes_per_s"], sh["shared_scan_128_linear_standard_codes_per_s"]), 1)
    else:
        line["cpu_baseline"] = None
    # cfg 4 at its stated size needs the GPU to itself: free the sweep's columns first
    if rank == 0 and world == 1 and not args.no_extras and args.logn == 30 and not args.no_cfg4:
        keep9 = columns[9]
        columns.clear()
        del bitvec
        torch.cuda.empty_cache()
        extras_out["cfg4"] = measure_cfg4(torch, api, keep9, n, peak)
        summary["linear_k1024_w9_cfg4"] = {k: extras_out["cfg4"][k] for k in ("log2_n", "chunks", "ms", "gcodes_per_s", "frac")}
    if rank == 0:
        extras_file = args.extras_file or os.path.join("gpurun_out", f"bench_extras_N{world}.json")
        try:
            os.makedirs(os.path.dirname(os.path.join(ROOT, extras_file)) or ".", exist_ok=True)
            with open(os.path.join(ROOT, extras_file), "w") as f:
                json.dump({"line": line, **extras_out}, f)
            line["extras_file"] = extras_file
        except OSError as e:
            line["extras_file"] = f"not written: {e}"
        emit(line)
    D.close()
    return 0


def shared_point(D, api, columns, n, kind, w, K, peak):
    """One (op, w, K) cell of the metric's table on this run's GPUs: device-resident, best of 5 event
    pairs, max over ranks."""
    torch = D.torch
    lo_d, hi_d = random_ranges(torch, w, K, seed=42 + K)
    cnt = torch.empty(K, dtype=torch.int64, device="cuda")
    if kind == "shared":
        buf = torch.empty((K, (api.bitvec_bytes(n) + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
        fn = lambda: api.shared_scan(columns[w], n, w, lo_d, hi_d, out=buf, counts=cnt)  # noqa: E731
    else:
        buf = torch.empty((4 * ((n + 31) // 32), K), dtype=torch.uint8, device="cuda")
        fn = lambda: api.shared_scan_linear(columns[w], n, w, lo_d, hi_d, out=buf, counts=cnt)  # noqa: E731
    ms = D.max_over_ranks(launch_ms(fn))
    b = n * (w / 8 + K / 8)
    del buf
    return {"ms": round(ms, 4), "gcodes_per_s": round(n / ms / 1e6, 1), "gb_per_s": round(b / ms / 1e6, 1),
            "frac": round(b / ms / 1e6 / peak, 4)}


def measure_cfg4(torch, api, column9, n, peak):
    """BASELINE configs[3] at its stated size: linear shared scan, K = 1024 equality+range predicates over
    2^30 9-bit codes = 128 GiB of output.  Resident when the GPU has room for it, else in `chunks` equal
    column chunks that reuse one output region (stated in the result)."""
    w, K = 9, 1024
    lo_d, hi_d = random_ranges(torch, w, K, seed=42 + K)
    cnt = torch.empty(K, dtype=torch.int64, device="cuda")
    free, _total = torch.cuda.mem_get_info()
    chunks = 1
    while chunks < 64 and (n // chunks) // 8 * K + (2 << 30) > free:
        chunks *= 2
    nc = n // chunks
    out = torch.empty(((nc // 8), K), dtype=torch.uint8, device="cuda")
    bytes_per_chunk_in = nc * w // 8  # a multiple of 16 bytes for these sizes

    def run():
        for c in range(chunks):
            api.shared_scan_linear(column9[c * bytes_per_chunk_in:], nc, w, lo_d, hi_d, out=out, counts=cnt)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        run()
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / reps
    nbytes = n * (w / 8 + K / 8)
    # the big run's first rows equal a small run of the same kernel family on the same column prefix
    # (which the parity tests pin against the oracle)
    ns = 1 << 19
    small, _ = api.shared_scan_linear(column9[(chunks - 1) * bytes_per_chunk_in:], ns, w, lo_d, hi_d)
    assert torch.equal(small[: ns // 8], out[: ns // 8]), "cfg4: full-size rows != small-run rows"
    assert int(cnt.sum().item()) > 0
    del small
    res = {"op": "linear", "w": w, "K": K, "log2_n": int(math.log2(n)), "chunks": chunks,
           "out_gib": round(K * (n // 8) / 2**30, 1), "resident_out_gib": round(K * (nc // 8) / 2**30, 1),
           "ms": round(ms, 3), "gcodes_per_s": round(n / ms / 1e6, 2), "gb_per_s": round(nbytes / ms / 1e6, 1),
           "frac": round(nbytes / ms / 1e6 / peak, 4)}
    del out
    return res


def measure_cfg1(torch, api, peak):
    """BASELINE configs[0] on the GPU: single range-predicate scan over 2^24 11-bit codes -- a 3.8 us
    kernel at the HBM roofline, so it is timed as 100 launches captured in ONE CUDA graph (launch
    latency amortised), and as 100 plain stream launches for comparison."""
    w, n = 11, 1 << 24
    col = api.generate_packed(n, w, seed=42 + w)
    bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    lo, hi = predicate(w)
    launches = 100
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            api.scan(col, n, w, lo, hi, out=bv, count=cnt)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(launches):
                api.scan(col, n, w, lo, hi, out=bv, count=cnt)
        for _ in range(2):
            g.replay()
        s.synchronize()
        best_graph = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            g.replay()
            b.record(s)
            b.synchronize()
            best_graph = min(best_graph, a.elapsed_time(b) / launches)
        best_plain = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            for _ in range(launches):
                api.scan(col, n, w, lo, hi, out=bv, count=cnt)
            b.record(s)
            b.synchronize()
            best_plain = min(best_plain, a.elapsed_time(b) / launches)
    nbytes = n * (w / 8 + 1 / 8)
    ideal_us = nbytes / (peak * 1e9) * 1e6
    return {"w": w, "log2_n": 24, "launches_per_graph": launches, "us_per_scan_graph": round(best_graph * 1e3, 2),
            "us_per_scan_stream": round(best_plain * 1e3, 2), "ideal_us_at_measured_peak": round(ideal_us, 2),
            "pct_of_ideal_graph": round(100 * ideal_us / (best_graph * 1e3), 1),
            "gcodes_per_s_graph": round(n / best_graph / 1e6, 1),
            "note": "25 MB working set: L2-resident after the first launch, so the fraction can exceed the HBM roofline"}


def measure_e2e_pageable(torch, api, columns, n):
    """What the C++ drop-in's callers really hand in: PAGEABLE host memory (make_unique / std::vector,
    src/simd_scan.hpp:46-120).  The host entry points stage such buffers through pinned bounce buffers
    with a few copy threads (capi.cu HostPipe); numbers for the scan (input-bound) and the decompression
    (output-bound) at 17 bit, next to the same calls on pinned buffers."""
    import numpy as np
    w = 17
    rows = []
    nb = api.packed_bytes(n, w)
    pageable_in = np.empty(nb, dtype=np.uint8)
    pinned_in = api.pinned_empty(nb)
    pinned_in.copy_(columns[w][:nb])
    pageable_in[:] = pinned_in.numpy()
    m = (1 << w) - 1
    n2 = min(n, 1 << 28)
    for kind, buf in (("pageable", pageable_in), ("pinned", pinned_in)):
        out = np.zeros(api.bitvec_bytes(n), dtype=np.uint8) if kind == "pageable" else api.pinned_empty(api.bitvec_bytes(n))
        api.host_scan(buf, n, w, m // 4, m // 2, tail_granule=1, out=out)
        t0 = time.perf_counter()
        for _ in range(3):
            api.host_scan(buf, n, w, m // 4, m // 2, tail_granule=1, out=out)
        sec = (time.perf_counter() - t0) / 3
        rows.append({"op": "scan", "w": w, "host_memory": kind, "gcodes_per_s": round(n / sec / 1e9, 2),
                     "h2d_gb_per_s": round(n * w / 8 / sec / 1e9, 1)})
        dec = np.empty(n2, dtype=np.int32) if kind == "pageable" else api.pinned_empty(n2 * 4).view(torch.int32)
        api.host_decompress(buf, n2, w, out=dec)
        t0 = time.perf_counter()
        for _ in range(3):
            api.host_decompress(buf, n2, w, out=dec)
        sec = (time.perf_counter() - t0) / 3
        rows.append({"op": "decompress", "w": w, "host_memory": kind, "gcodes_per_s": round(n2 / sec / 1e9, 2),
                     "d2h_gb_per_s": round(n2 * 4 / sec / 1e9, 1)})
    return rows


def extras(args, api, columns, n, peak):
    """Per-K shared-scan numbers and the decompression sweep (BASELINE.json configs 2-4), measured
    after the timed steps with the same event discipline (3 warm-ups, mean of 5)."""
    import torch

    def timed(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    out = {"shared": [], "unpack": [], "pack": [], "positions": [], "scan_positions": [], "calibration": calibration(api, timed, peak),
           "cfg1": measure_cfg1(torch, api, peak)}
    out["e2e_pageable"] = measure_e2e_pageable(torch, api, columns, n)
    if not args.no_sweep:
        out["shared_sweep"] = shared_sweep(args, api, columns, n, peak)
    for kind, w, K, logn in (("shared", 17, 8, 30), ("shared", 17, 64, 30), ("shared", 23, 64, 30),
                             ("linear", 9, 64, 30), ("linear", 9, 1024, 27)):
        nn = 1 << min(logn, args.logn)
        lo_d, hi_d = random_ranges(torch, w, K, seed=42)
        cnt = torch.empty(K, dtype=torch.int64, device="cuda")
        if kind == "shared":
            buf = torch.empty((K, (api.bitvec_bytes(nn) + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
            ms = timed(lambda: api.shared_scan(columns[w], nn, w, lo_d, hi_d, out=buf, counts=cnt))
        else:
            buf = torch.empty((4 * ((nn + 31) // 32), K), dtype=torch.uint8, device="cuda")
            ms = timed(lambda: api.shared_scan_linear(columns[w], nn, w, lo_d, hi_d, out=buf, counts=cnt))
        b = nn * (w / 8 + K / 8)
        out["shared"].append({"op": kind, "w": w, "K": K, "n": nn, "ms": round(ms, 3),
                              "gcodes_per_s": round(nn / ms / 1e6, 2), "gb_per_s": round(b / ms / 1e6, 1),
                              "frac_of_measured_peak": round(b / ms / 1e6 / peak, 4)})
        del buf
    # ascending position list from the scan's bitvector, at three selectivities (w = 17); algorithmic
    # bytes = ONE read of the bitvector + 8 bytes per match
    w = 17
    m = (1 << w) - 1
    bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    cnt1 = torch.empty(1, dtype=torch.int64, device="cuda")
    for sel in (0.001, 0.01, 0.25):
        hi_v = max(0, int(sel * (m + 1)) - 1)
        api.scan(columns[w], n, w, 0, hi_v, out=bv, count=cnt1)
        hits = int(cnt1.item())
        pos_buf = torch.empty(max(hits, 1), dtype=torch.int64, device="cuda")
        num_buf = torch.empty(1, dtype=torch.int64, device="cuda")
        api.positions(bv, n, capacity=hits, out=pos_buf, num=num_buf)  # sizes the context workspace (untimed)
        ms = timed(lambda: api.positions(bv, n, capacity=hits, out=pos_buf, num=num_buf), reps=5, warm=2)
        assert int(num_buf.item()) == hits
        b = n / 8 + 8.0 * hits
        out["positions"].append({"w": w, "n": n, "selectivity": sel, "hits": hits, "ms": round(ms, 4),
                                 "gcodes_per_s": round(n / ms / 1e6, 1), "gb_per_s": round(b / ms / 1e6, 1),
                                 "frac_of_measured_peak": round(b / ms / 1e6 / peak, 4)})
        # the north star's scan output in ONE call (sss_scan_range_positions): the scan fills the position
        # list's segment counts itself, so the list costs group offsets + expansion, no count pass;
        # bytes = packed column + bitvector written + 8 per match (the expansion's re-read is not counted)
        bv2 = torch.empty_like(bv)
        fused = lambda: api.scan_positions(columns[w], n, w, 0, hi_v, capacity=hits, out=bv2, pos_out=pos_buf,  # noqa: E731
                                           count=cnt1)
        ms_f = timed(fused, reps=5, warm=2)
        ms_s = timed(lambda: api.scan(columns[w], n, w, 0, hi_v, out=bv2, count=cnt1), reps=5, warm=2)
        assert int(cnt1.item()) == hits
        bf = n * (w / 8 + 1 / 8) + 8.0 * hits
        out["scan_positions"].append({"w": w, "n": n, "selectivity": sel, "hits": hits, "ms": round(ms_f, 4),
                                      "ms_scan_alone": round(ms_s, 4), "ms_two_calls": round(ms_s + ms, 4),
                                      "gcodes_per_s": round(n / ms_f / 1e6, 1), "gb_per_s": round(bf / ms_f / 1e6, 1),
                                      "frac_of_measured_peak": round(bf / ms_f / 1e6 / peak, 4)})
        del pos_buf, bv2
    del bv
    n2 = 1 << min(28, args.logn)
    dec = torch.empty(n2, dtype=torch.int32, device="cuda")
    # the compressor: int32 codes -> packed stream, 4 + w/8 bytes per code
    for w in (1, 9, 17, 32):
        api.decompress(columns[w], n2, w, out=dec)
        repacked = torch.empty(api.packed_bytes(n2, w), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: api.compress(dec, w, out=repacked), reps=3, warm=2)
        assert torch.equal(repacked[: n2 * w // 8], columns[w][: n2 * w // 8]), "pack(unpack(x)) != x"
        b = n2 * (4 + w / 8)
        out["pack"].append({"w": w, "n": n2, "ms": round(ms, 4), "gcodes_per_s": round(n2 / ms / 1e6, 1),
                            "gb_per_s": round(b / ms / 1e6, 1), "frac_of_measured_peak": round(b / ms / 1e6 / peak, 4)})
        del repacked
    for w in WIDTHS:
        ms = timed(lambda: api.decompress(columns[w], n2, w, out=dec), reps=3, warm=2)
        b = n2 * (w / 8 + 4)
        out["unpack"].append({"w": w, "n": n2, "ms": round(ms, 4), "gcodes_per_s": round(n2 / ms / 1e6, 1),
                              "gb_per_s": round(b / ms / 1e6, 1), "frac_of_measured_peak": round(b / ms / 1e6 / peak, 4)})
    return out


def calibration(api, timed, peak):
    """HBM bandwidth for the access mixes of this path (the reference's
    bench_memory / bench_memcpy, src/benchmark_misc.cpp:9-52): streaming read, write, copy, and
    read:write = W:1 (the traffic shape of the fused scan at W bits) over 4 GiB.  GB/s counts the bytes
    actually moved; `vs_copy_peak` is against MEASURED_PEAKS.json, the roofline denominator."""
    import torch

    nbytes = 4 << 30
    a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    a.random_(0, 256)
    rows = []
    for mode, ratio in (("read", 1), ("write", 1), ("copy", 1), ("mix", 1), ("mix", 2), ("mix", 4), ("mix", 8),
                        ("mix", 16), ("mix", 32)):
        ms = timed(lambda: api.calibrate(mode, a, b, nbytes, ratio), reps=5, warm=2)
        moved = {"read": nbytes, "write": nbytes, "copy": 2 * nbytes}.get(mode, nbytes + nbytes // ratio)
        rows.append({"mode": mode if mode != "mix" else f"read:write = {ratio}:1", "gb_per_s": round(moved / ms / 1e6, 1),
                     "vs_copy_peak": round(moved / ms / 1e6 / peak, 4)})
    # the same mixes at the SIZE of the 2^30-code scans at 1-4 bit (128-512 MiB in, 128 MiB out): a 45-100 us
    # launch pays its ramp and drain, so this -- not the 4 GiB figure -- is what those widths can be held to
    for w in (1, 2, 3, 4):
        nb = w << 27
        ms = launch_ms(lambda: api.calibrate("mix", a, b, nb, w))
        moved = nb + nb // w
        rows.append({"mode": f"read:write = {w}:1 at the {w}-bit scan's size ({nb >> 20} MiB in)", "us": round(ms * 1e3, 1),
                     "gb_per_s": round(moved / ms / 1e6, 1), "vs_copy_peak": round(moved / ms / 1e6 / peak, 4)})
    del a, b
    return rows


def launch_ms(fn, warm=2, rounds=5):
    """Duration of one launch of `fn` in ms: best of `rounds` event pairs, each around enough back-to-back
    launches to fill about a millisecond (an event pair around ONE 50 us launch reads ~10 us too long)."""
    import torch

    def timed_once(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    reps = max(1, min(10, int(1.0 / max(timed_once(1), 1e-3))))
    return min(timed_once(reps) for _ in range(rounds))


def shared_sweep(args, api, columns, n, peak):
    """The metric's table: shared scan (K bitvectors + K counts) and linear-layout shared scan at EVERY
    bit width 1..32 and K in {1, 8, 64, 1024}, device-resident (same method as tools/sweep_wk.py: K
    uniform-random ranges, 2 warm-ups, best of 5 event pairs, each around up to 10 back-to-back launches).  n is halved until the output fits a
    16 GiB arena (K = 1024: 2^27 codes).  Rows are [op, w, K, log2 n, ms, G codes/s, GB/s, fraction of
    the measured HBM peak]; bytes = n * (w/8 + K/8)."""
    import math

    import torch

    max_out = 16 << 30
    arena = torch.empty(max_out + (1 << 20), dtype=torch.uint8, device="cuda")
    rows = []
    for w in WIDTHS:
        m = (1 << w) - 1
        for K in (1, 8, 64, 1024):
            n2 = n
            while K * (n2 // 8) > max_out:
                n2 //= 2
            g = torch.Generator().manual_seed(42 + K)
            lo = torch.randint(0, m + 1, (K,), generator=g)
            hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
            lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
            cnt = torch.empty(K, dtype=torch.int64, device="cuda")
            for op in ("shared", "linear"):
                if op == "shared":
                    stride = (api.bitvec_bytes(n2) + 127) // 128 * 128
                    buf = arena[: K * stride].view(K, stride)
                    fn = lambda: api.shared_scan(columns[w], n2, w, lo_d, hi_d, out=buf, counts=cnt)  # noqa: E731
                else:
                    nrows = 4 * ((n2 + 31) // 32)
                    buf = arena[: nrows * K].view(nrows, K)
                    fn = lambda: api.shared_scan_linear(columns[w], n2, w, lo_d, hi_d, out=buf, counts=cnt)  # noqa: E731
                best = launch_ms(fn)
                nbytes = n2 * (w / 8 + K / 8)
                rows.append([op, w, K, int(math.log2(n2)), round(best, 4), round(n2 / best / 1e6, 1),
                             round(nbytes / best / 1e6, 1), round(nbytes / best / 1e6 / peak, 4)])
    del arena
    return {"columns": ["op", "w", "K", "log2_n", "ms", "gcodes_per_s", "gb_per_s", "frac_of_measured_peak"],
            "rows": rows}


def run_cfg5(args):
    """`--workload cfg5`: BASELINE configs[4] alone (also part of every default run's `summary`)."""
    from shared_simd_scan_b200 import api
    D = Dist()
    peak, _ = measured_peak()
    res = measure_cfg5(D, api, args.cfg5_logn, args.steps, peak)
    if D.rank == 0:
        emit({"workload": f"cfg5: shared scan K=64 over {D.world} x 2^{args.cfg5_logn} 23-bit codes, range-sharded, "
                          "NCCL all-reduce of the 64 counts", "n_gpus": D.world, "steps": args.steps,
              "value": res["gcodes_per_s"] * 1e9, "unit": "codes/s", "scaling": "weak", **res})
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--logn", type=int, default=30, help="log2 codes per width per GPU (BASELINE: 30)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-logn", type=int, default=28, help="cpu_baseline sample: log2 codes per width")
    ap.add_argument("--ref-logn", type=int, default=28, help="--impl reference sample: log2 codes per width per step")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "cfg5"],
                    help="sweep = the contract line; cfg5 = the sharded shared scan (K=64, 23 bit) alone")
    ap.add_argument("--cfg5-logn", type=int, default=31, help="cfg5: log2 codes per GPU")
    ap.add_argument("--extras-file", default=None, help="side file for the full tables (default gpurun_out/bench_extras_N<gpus>.json)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the (w x K) shared-scan table of the extras")
    ap.add_argument("--no-cfg4", action="store_true", help="skip cfg 4 at its stated size (128 GiB of output)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner,
    # compiler chatter) is sent to stderr instead, and the line is written to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    global _json_out
    _json_out = os.fdopen(saved, "w")
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "cfg5":
        return run_cfg5(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
