"""bench.py's JSON contract (driver-facing): the reference arm runs on the CPU here and must print one line with the
keys the driver reads; the GPU arm's line is checked for the same keys plus roofline / cpu_baseline / aux on a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["unit"] == "rollouts/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--steps", "2", "--warmup", "3", "--no-cpu-baseline"], 900)
    assert BASE_KEYS <= set(d) and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["gpu_launches"] > 0 and d["value"] > 1e6 and d["vs_baseline"] is None and d["scaling"] == "weak"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.02 and e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert "error" not in d["aux"] and d["aux"]["coef_draws_mt19937_ms"] > 0
