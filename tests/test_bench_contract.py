"""CPU: the reference arm of bench.py prints exactly one JSON line with the keys the measurement contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                      # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scans/sec" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] >= 3 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - 1.0) < 1e-6
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # "reference" = the reference's own sources from oracle/_ref (built where /root/reference exists, shipped with the snapshot),
    # "port" = the oracle restatement where they are not available
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_odometry.so"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] == 2 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
