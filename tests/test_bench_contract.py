"""bench.py contract checks that need no GPU: the reference arm runs the unmodified reference (oracle/_ref) on the
host cores and prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(R):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if "unavailable" in d:                      # only where oracle/_ref could not be built (no /root/reference)
        return
    assert d["metric"] == "converged Keff solves/sec" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"]


def test_bench_names_the_baseline_metric_and_workload():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "converged Keff solves/sec" in base["metric"] and 'METRIC, UNIT = "converged Keff solves/sec", "solves/s"' in src
    assert "10,000 synthetic 128" in base["configs"][1] and "N_IMG, NPIX = 10000, 128" in src
    for key in ('"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"', '"h2d_bytes_per_step"', '"traffic"'):
        assert key in src
