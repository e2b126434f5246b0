"""the benchmark CLI front-end prints the reference's report lines
(src/benchmark.cpp:14-36), so the reference's own result scraper (scripts/prepare_shared_scan_results.py:14-20,
restated below line for line in behaviour) turns a GPU run into the CSV its plot script reads."""
import csv
import io
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_output(data_size, predicate_count, output):
    """What scripts/prepare_shared_scan_results.py does with the CLI's stdout: lines starting with '*',
    variant = text between '* ' and ': ', average = text between ': ' and '; ' minus the unit."""
    rows = []
    for line in output.splitlines():
        if not line.startswith('*'):
            continue
        variant = line[2:].split(': ')[0]
        avg_runtime_ms = line.split(': ')[1].split('; ')[0][:-2]
        rows.append([data_size, predicate_count, variant, avg_runtime_ms])
    return rows


@pytest.mark.parametrize("K", [8, 64])
def test_cli_sharedscan_feeds_the_reference_scraper(K):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "shared_simd_scan_cli.py"), "40", "5", "sharedscan", str(K)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "first mismatch" not in r.stdout, r.stdout          # the CLI checks its results like check_scan_result
    assert "finished benchmark" in r.stdout
    assert f"predicate key count: {K}" in r.stdout
    rows = parse_output(40, K, r.stdout)
    variants = [row[2] for row in rows]
    # the two variants scripts/plot_shared_scan_results.py compares, plus the linear layout
    assert variants == ["sse 128, sequential (unrolled)", "sse 128, standard", "sse 128, linear, standard"]
    buf = io.StringIO()
    wr = csv.writer(buf)
    wr.writerow(['data_size', 'predicate_count', 'variant', 'avg_runtime_ms'])
    wr.writerows(rows)
    back = list(csv.DictReader(io.StringIO(buf.getvalue())))
    for row in back:
        assert float(row["avg_runtime_ms"]) > 0 and int(row["predicate_count"]) == K
    # one pass for K predicates must beat K passes (best of the five repetitions each: a 0.1 ms measurement right
    # after start-up also sees clock ramps and first-launch costs, and back-to-back single scans overlap their
    # launches since the end of round 2)
    best = {}
    for line in r.stdout.splitlines():
        if line.startswith('*'):
            best[line[2:].split(': ')[0]] = min(float(x) for x in line.split('[')[1].split(']')[0].split(', '))
    assert best["sse 128, standard"] < best["sse 128, sequential (unrolled)"], r.stdout


@pytest.mark.parametrize("bench", ["scan", "decompression"])
def test_cli_other_benches(bench):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "shared_simd_scan_cli.py"), "40", "2", bench],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "first mismatch" not in r.stdout and "finished benchmark" in r.stdout
    rows = parse_output(40, 1, r.stdout)
    assert len(rows) == 1 and float(rows[0][3]) > 0
