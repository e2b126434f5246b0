"""bench.py's contract on a machine without a GPU: the reference arm prints ONE JSON line with the
contract's keys (oracle/_ref or the oracle port on the host cores, a tiny sample here), and the
product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-logn", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scanned_codes_per_s" and d["unit"] == "codes/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "codes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert len(lines[0]) < 6000, "the driver parses ONE compact line (round 1's 27 KB line came back unparsed)"


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT,
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "1", "--logn", "16", "--no-extras", "--no-cpu")
    assert r.returncode != 0
    assert r.stdout.strip() == ""          # no JSON line from a CPU fallback
    assert "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_product_arm_prints_one_compact_parsable_line(tmp_path):
    """The product arm on a GPU (small column so that it runs in seconds): ONE JSON line, < 6 KB, the same
    keys the driver reads at every N, the full tables in the side file it names."""
    extras = tmp_path / "extras.json"
    r = _run("--steps", "2", "--warmup", "1", "--logn", "22", "--e2e-steps", "1", "--cpu-logn", "18", "--cfg5-logn", "22",
             "--no-sweep", "--extras-file", str(extras))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    assert len(lines[0]) < 6000
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "summary", "per_width_frac", "extras_file"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] == 2 * 32 and len(d["per_width_frac"]) == 32
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["frac"] > 0
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["summary"]["cfg5_shared_k64_w23"]["gcodes_per_s"] > 0
    side = json.loads(extras.read_text())
    for key in ("per_width", "shared", "unpack", "pack", "positions", "calibration", "cfg1"):
        assert key in side, key
