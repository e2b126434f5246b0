#!/usr/bin/env python
"""Front-end with the reference benchmark CLI's contract:

    shared_simd_scan_cli.py data_size repetitions bench_name [predicate_count]
    bench_name = decompression | scan | sharedscan      ("_" = default for the first two arguments)

Same positional arguments and defaults as the reference (src/main.cpp:12-73, src/benchmark.hpp:4-5:
500 MiB of packed 9-bit data, 5 repetitions, K = 8), same synthetic data (src/benchmark.cpp:81,173,
277: i & 511 / i % 5 with key 3 / i % K % 512 with keys 0..K-1) and the same report lines
(`* name: avg ms; [r1, r2, ...] ms`, src/benchmark.cpp:14-36), so the reference's
scripts/prepare_shared_scan_results.py and plot_shared_scan_results.py work against the GPU build
unchanged.  Times are device times of the device-resident entry points (CUDA events), printed with
three decimals (the reference truncates to whole milliseconds).

The variant names the reference's plot script keys on are kept for the two it compares:
"sse 128, sequential (unrolled)" = K independent single scans (the column is re-read K times),
"sse 128, standard" = the shared scan kernel.  Every line is checked against the uncompressed input
like the reference's check_scan_result (src/benchmark.cpp:110-121).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

W = 9  # the reference's BITS_NEEDED (src/simd_scan.hpp:12)


def usage():
    print("Format: shared_simd_scan_cli.py data_size repetitions bench_name [bench_args...]")
    print("bench_name = decompression | scan | sharedscan [predicate_count]")
    return 1


def report(name, ms_list):
    avg = sum(ms_list) / len(ms_list)
    print(f"* {name}: {avg:.3f} ms; [{', '.join(f'{m:.3f}' for m in ms_list)}] ms", flush=True)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        out.append(a.elapsed_time(b))
    return out


def column(n, pattern):
    idx = torch.arange(n, dtype=torch.int64, device="cuda")
    codes = pattern(idx).to(torch.int32)
    return codes, api.compress(codes, W)


def check_bits(bitvec, truth, n):
    got = torch.from_numpy(np.unpackbits(bitvec[: (n + 7) // 8].cpu().numpy(), bitorder="little")[:n].astype(bool))
    bad = torch.nonzero(got != truth.cpu())
    if len(bad):
        print(f"first mismatch at index {int(bad[0])}")


def main(argv):
    if len(argv) < 4:
        return usage()
    data_size = (500 << 20) if argv[1] == "_" else int(argv[1]) << 20
    reps = 5 if argv[2] == "_" else int(argv[2])
    bench = argv[3]
    n = data_size * 8 // W
    if bench == "decompression":
        codes, packed = column(n, lambda i: i & 511)
        print("## decompression benchmarks ##")
        print(f"compressed input: {n} ({data_size} bytes)")
        out = torch.empty(n, dtype=torch.int32, device="cuda")
        report("b200 unpack", timed(lambda: api.decompress(packed, n, W, out=out), reps))
        if not torch.equal(out, codes):
            print("first mismatch at index", int(torch.nonzero(out != codes)[0]))
    elif bench == "scan":
        codes, packed = column(n, lambda i: i % 5)
        print("## scan benchmarks ##")
        print(f"compressed input: {n} ({data_size} bytes)")
        bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
        cnt = torch.empty(1, dtype=torch.int64, device="cuda")
        report("b200 fused unpack+scan", timed(lambda: api.scan(packed, n, W, 3, out=bv, count=cnt), reps))
        check_bits(bv, codes == 3, n)
    elif bench == "sharedscan":
        K = int(argv[4]) if len(argv) > 4 else 8
        codes, packed = column(n, lambda i: i % K % 512)
        keys = list(range(K))
        print("## shared scan benchmarks ##")
        print("# B200 (sm_100a) kernels of shared_simd_scan_b200; the variant labels below are the reference's, kept so")
        print("# that scripts/prepare_shared_scan_results.py / plot_shared_scan_results.py work unchanged:")
        print("#   'sse 128, sequential (unrolled)' = K single-predicate GPU scans, 'sse 128, standard' = the shared-scan")
        print("#   kernel (K bitvectors), 'sse 128, linear, standard' = the linear-layout shared-scan kernel")
        print(f"compressed input: {n} ({data_size} bytes)")
        print(f"predicate key count: {K}")
        stride = (api.bitvec_bytes(n) + 15) // 16 * 16
        outs = torch.empty((K, stride), dtype=torch.uint8, device="cuda")
        cnts = torch.empty(K, dtype=torch.int64, device="cuda")
        cnt = torch.empty(1, dtype=torch.int64, device="cuda")

        def sequential():
            for k in range(K):
                api.scan(packed, n, W, keys[k], out=outs[k], count=cnt)

        lo = torch.tensor(keys, dtype=torch.int32, device="cuda")
        report("sse 128, sequential (unrolled)", timed(sequential, reps))
        report("sse 128, standard", timed(lambda: api.shared_scan(packed, n, W, lo, lo, out=outs, counts=cnts), reps))
        for k in (0, K - 1):
            check_bits(outs[k], codes == keys[k], n)
        lin = torch.empty((4 * ((n + 31) // 32), K), dtype=torch.uint8, device="cuda")
        report("sse 128, linear, standard",
               timed(lambda: api.shared_scan_linear(packed, n, W, lo, lo, out=lin, counts=cnts), reps))
    else:
        return usage()
    print("finished benchmark")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
