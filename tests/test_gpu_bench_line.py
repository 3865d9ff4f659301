"""GPU: `bench.py` on a small shape prints one JSON line carrying every key of the contract (value, e2e with the byte
counts, clocks, gpu_launches, roofline with per-launch fields, cpu_baseline)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_bench_line_contract_keys():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--shape", "120x320x64"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2 * (120 * 320 * 3 + 120 * 320) * 2 and e["d2h_bytes_per_step"] == 120 * 320 * 4 * 2
    assert e["value"] != d["value"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["avg_launch_ms"] > 0 and rf["algorithmic_bytes_per_launch"] > 0 and "traffic" in rf
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and cb["sample"]
    assert d["reference_order_mode"]["value"] > 0
