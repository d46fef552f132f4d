"""CPU: the bench line committed under profiles/ carries every key of the bench.py contract (metric, whole-job value, e2e through host
buffers, roofline of the dominant kernel against a measured peak, CPU baseline, clocks), with consistent arithmetic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line():
    return json.load(open(os.path.join(ROOT, "profiles", "r1_bench_final.json")))


def test_contract_keys_and_arithmetic():
    d = _line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = env transitions / s: 4096 envs x <= 25 steps per iteration
    per_step = d["value"] * d["ms_per_step"] / 1e3
    assert 0.5 * 4096 * 25 <= per_step <= 4096 * 25
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0.5 * d["value"] < e["value"] <= 1.05 * d["value"]
    assert d["gpu_launches"] > d["steps"] * 4096
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and r["traffic"] is not None
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["flop_per_launch"] / (r["launch_us"] * 1e-6) / 1e12) < 1e-6 * r["achieved"]
    b = d["cpu_baseline"]
    assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_bench_cli_parses():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "--impl" in out.stdout and "--collective" in out.stdout
