"""bench.py prints exactly one JSON line on stdout with the keys the driver reads — both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold the one result line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-seconds", "0.05")
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "Mpx/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("cfg2")
    import bench
    assert d["config"] == bench.bench_config("recursive")       # the same workload description as our arm's line
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "band" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--steps", "16", "--warmup", "3", "--no-cpu")
    assert BASE_KEYS | {"clocks", "roofline"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["steps"] == 16 and d["scaling"] == "weak"
    import bench
    assert d["config"] == bench.bench_config("recursive")
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 5 * 16          # pyramid x2, rows, columns, finalize per step
    assert d["value"] > 1000 and d["e2e"]["value"] > 100
    assert d["e2e"]["h2d_bytes_per_step"] == 3840 * 2160 * 9 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    assert rf["kernel"] in ("k_iir_cols", "k_iir_rows") and rf["traffic"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert abs(d["score_check"] - 82.814139) < 1e-3   # the bench pair as scored in round 1 (oracle parity: test_gpu_parity)
