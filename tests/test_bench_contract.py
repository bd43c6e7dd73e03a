"""CPU: the JSON contract of the reference arm of bench.py (`--impl reference`), which times the unmodified reference
(baseline/_ref, installed by baseline/install_reference.sh) -- or, when that directory is absent, the float64 oracle
port -- on the host cores and needs no GPU: the keys the driver reads, the tier's `cpu_baseline` and `e2e` objects."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "cfg3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "KL-NMF iterations/sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000.0
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "cfg3" in d["config"]["workload"] and d["config"]["k"] == 256
    cb = d["cpu_baseline"]
    installed = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "multimodal"))
    assert cb["kind"] == ("reference" if installed else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "extrapolated" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
