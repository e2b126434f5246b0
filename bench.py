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
Extra keys (per_width, shared, unpack) carry the per-bit-width and per-K numbers the metric asks for.
"""
from __future__ import annotations

import argparse
import json
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
    (profiles/r01_scan_traffic.json; tools/scan_traffic.py + tools/summarise_traffic.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_scan_traffic.json")) as f:
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

def cpu_reference_sweep(logn, reps, threads):
    """The reference's own CPU scan over a bounded sample: 2^logn uniform codes per width.

    oracle/_ref (the unmodified reference compiled per width) when present: scan_256_unrolled for
    w <= 16 (the "reference AVX2 scan"), scan_128_unrolled above (the AVX path is only exact up to
    16 bits); every host thread scans its own range of the column.  If
    oracle/_ref is absent the plain-C oracle port is timed instead (kind "port").
    Returns (codes_per_s over the whole sweep, per-width seconds, kind)."""
    import numpy as np
    from oracle import binding as B
    n = 1 << logn
    rng = np.random.default_rng(42)
    per_w = {}
    kind = "reference" if all(B.ref_available(w) for w in WIDTHS) else "port"
    for w in WIDTHS:
        nbytes = B.compressed_buffer_size(w, n)
        packed = np.zeros(nbytes, dtype=np.uint8)
        payload = n * w // 8
        packed[:payload] = np.frombuffer(rng.bytes(payload), dtype=np.uint8)  # uniform bits == uniform codes
        key = 3 & ((1 << w) - 1)  # the reference only has equality predicates (src/simd_scan.cpp:47)
        if kind == "reference":
            variant = B.SCAN["256_unrolled"] if w <= 16 else B.SCAN["128_unrolled"]
            import ctypes as C
            hits = C.c_longlong(0)
            s = B.ref(w).ref_time_scan(variant, key, packed.ctypes.data, n, threads, reps, C.byref(hits))
        else:
            s, _ = B.time_scan_range(packed, n, w, key, key, threads, reps)
        per_w[w] = s
    total = sum(per_w.values())
    return len(WIDTHS) * n / total, per_w, kind


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
    # warm-up + timed steps: each step = the sweep over the bounded sample
    for _ in range(args.warmup):
        cpu_reference_sweep(min(logn, 22), 1, threads)
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

def run_cfg5(args):
    """Optional workload (not the contract line): BASELINE configs[4] -- shared scan, K = 64 range
    predicates over 23-bit codes, the column range-sharded over the ranks (2^31 codes per GPU = 16 Gi
    codes on 8 GPUs), per-query counts combined with one NCCL all-reduce.  `--workload cfg5`."""
    import torch
    import torch.distributed as dist

    from shared_simd_scan_b200 import api
    from shared_simd_scan_b200 import dist as sdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, K = 23, 64
    n_per = 1 << args.cfg5_logn
    shard = sdist.generate_shard(n_per * world, w, seed=42 + w)
    m = (1 << w) - 1
    g = torch.Generator().manual_seed(42)
    lo = torch.randint(0, m + 1, (K,), generator=g)
    hi = torch.clamp(lo + torch.randint(0, 2 * (m + 1) // K, (K,), generator=g), max=m)
    lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
    out = torch.empty((K, (api.bitvec_bytes(shard.count) + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
    counts = torch.empty(K, dtype=torch.int64, device="cuda")

    def step():
        api.shared_scan(shard.packed, shard.count, w, lo_d, hi_d, out=out, counts=counts)
        if world > 1:
            dist.all_reduce(counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    peak, _ = measured_peak()
    bytes_per_gpu = shard.count * (w / 8 + K / 8)
    if rank == 0:
        emit({
            "workload": f"cfg5: shared scan K=64 over {world} x 2^{args.cfg5_logn} 23-bit codes, range-sharded, "
                        "NCCL all-reduce of the 64 counts", "n_gpus": world, "steps": args.steps, "ms_per_step": ms,
            "value": world * shard.count / (ms * 1e-3), "unit": "codes/s", "scaling": "weak",
            "per_gpu_gb_per_s": round(bytes_per_gpu / ms / 1e6, 1),
            "per_gpu_frac_of_measured_peak": round(bytes_per_gpu / ms / 1e6 / peak, 4),
            "global_count_of_query_0": int(counts[0].item())})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from shared_simd_scan_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: shared_simd_scan_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
        if world > 1:
            dist.all_reduce(counts)  # the only collective on the path: 32 x int64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    t_begin.record()
    for k in range(args.steps):
        step()
    t_end.record()
    barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop(wall0, wall1)
    dev_ms = t_begin.elapsed_time(t_end)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = world * len(WIDTHS) * n / (ms_per_step * 1e-3)

    # roofline of the step's kernel (32 launches of scan_range_kernel<W>): algorithmic bytes of a step over the
    # device time of a step, both from the timed region above (so launch gaps count against the kernel)
    total_bytes = sum(n * (w / 8 + 1 / 8) for w in WIDTHS)
    achieved = total_bytes / (ms_per_step * 1e-3) / 1e9
    final_counts = counts.cpu().tolist()

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

    # ---- e2e: the host-buffer C-ABI call (what include/simd_scan.hpp wraps), pinned host memory ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    max_bytes = api.packed_bytes(n, 32)
    h_in = api.pinned_empty(max_bytes)
    h_out = api.pinned_empty(api.bitvec_bytes(n))
    e2e_sec, h2d, d2h = [], 0, 0
    e2e_counts = []
    for s in range(e2e_steps + 1):  # first pass = warm-up (staging buffers get allocated)
        sec = 0.0
        h2d = d2h = 0
        for w in WIDTHS:
            nb = api.packed_bytes(n, w)
            h_in[:nb].copy_(columns[w][:nb])  # refill the host column (untimed)
            torch.cuda.synchronize()
            lo, hi = predicate(w)
            t0 = time.perf_counter()
            _, c = api.host_scan(h_in, n, w, lo, hi, tail_granule=1, out=h_out)
            sec += time.perf_counter() - t0
            h2d += (n // 32) * w * 4
            d2h += n // 8 + 8
            if s == e2e_steps:
                e2e_counts.append(c)
        if s > 0:
            e2e_sec.append(sec)
    te = torch.tensor([statistics.mean(e2e_sec)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * len(WIDTHS) * n / float(te.item())
    del h_in, h_out
    if world == 1:
        assert e2e_counts == final_counts, "host-buffer path and device-resident path disagree on the match counts"

    line = {
        "metric": "scanned_codes_per_s", "value": value, "unit": "codes/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "codes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "sss_host_scan_range (pinned host buffers, 64 MiB chunks, 3 streams)"},
        "gpu_launches": args.steps * len(WIDTHS),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4),
                     "traffic": captured_traffic()[0] if args.logn == 30 else None,
                     "algorithmic_bytes": total_bytes,
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 32 launches of "
                                       "one step at 2^30 codes (profiles/r01_scan_traffic.json)",
                     "peak_source": peak_src,
                     "kernel": "sss::scan_range_kernel<W> (W = 1: scan_direct_kernel), the 32 launches of a step: "
                               "bytes = sum of n*(w/8+1/8), time = the step's device time in the timed region "
                               "(per-width launch durations in per_width: 10 back-to-back launches per event pair)",
                     "frac_of_8TBps_nominal": round(achieved / 8000.0, 4)},
        "per_width": per_width,
    }

    if rank == 0 and world == 1 and not args.no_extras:
        line.update(extras(args, api, columns, n, peak))
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        v, per_w, kind = cpu_reference_sweep(args.cpu_logn, 2, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": "codes/s", "cores": threads, "kind": kind,
            "sample": f"same sweep on 2^{args.cpu_logn} uniform codes per width (best of 2), equality key 3; "
                      "scan_256_unrolled (w<=16) / scan_128_unrolled (w>16) of the unmodified reference, column "
                      f"range-partitioned over {threads} host threads; {time.perf_counter() - t0:.1f} s of wall time"}
        line["cpu_reference_detail"] = cpu_reference_detail(threads)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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

    out = {"shared": [], "unpack": [], "pack": [], "positions": [], "calibration": calibration(api, timed, peak)}
    if not args.no_sweep:
        out["shared_sweep"] = shared_sweep(args, api, columns, n, peak)
    g = torch.Generator().manual_seed(42)
    for kind, w, K, logn in (("shared", 17, 8, 30), ("shared", 17, 64, 30), ("shared", 23, 64, 30),
                             ("linear", 9, 64, 30), ("linear", 9, 1024, 27)):
        nn = 1 << min(logn, args.logn)
        m = (1 << w) - 1
        lo = torch.randint(0, m + 1, (K,), generator=g)
        hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
        lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
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
    # ascending position list from the scan's bitvector, at three selectivities (w = 17)
    w = 17
    m = (1 << w) - 1
    bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    cnt1 = torch.empty(1, dtype=torch.int64, device="cuda")
    for sel in (0.001, 0.01, 0.25):
        hi_v = max(0, int(sel * (m + 1)) - 1)
        api.scan(columns[w], n, w, 0, hi_v, out=bv, count=cnt1)
        hits = int(cnt1.item())
        api.positions(bv, n, capacity=hits)  # sizes the context workspace outside the timed region
        ms = timed(lambda: api.positions(bv, n, capacity=hits), reps=3, warm=1)
        b = 2 * n / 8 + 8.0 * hits
        out["positions"].append({"w": w, "n": n, "selectivity": sel, "hits": hits, "ms": round(ms, 4),
                                 "gcodes_per_s": round(n / ms / 1e6, 1), "gb_per_s": round(b / ms / 1e6, 1),
                                 "frac_of_measured_peak": round(b / ms / 1e6 / peak, 4)})
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--logn", type=int, default=30, help="log2 codes per width per GPU (BASELINE: 30)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-logn", type=int, default=27, help="cpu_baseline sample: log2 codes per width")
    ap.add_argument("--ref-logn", type=int, default=26, help="--impl reference sample: log2 codes per width per step")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "cfg5"],
                    help="sweep = the contract line; cfg5 = optional sharded shared scan (K=64, 23 bit)")
    ap.add_argument("--cfg5-logn", type=int, default=31, help="cfg5: log2 codes per GPU")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the (w x K) shared-scan table of the extras")
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
