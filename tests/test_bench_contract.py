"""The driver's contract on bench.py: one JSON line on stdout with the agreed keys — the reference arm
on the CPU (runs here), our arm on a GPU (a short stream, every section switched on)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run_bench(*args, timeout=900):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("reassigned frames/sec") and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--seconds", "120", "--steps", "3", "--warmup", "3", "--cpu-seconds", "2", "--stream-pushes", "300",
                  "--batch-clips", "128", "--batch-steps", "1")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "pipeline_u8", "worst_case", "batch", "stream_latency",
                        "nfft_sweep", "e2e_display_rows"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["gpu_launches"] == d["steps"]                       # one kernel launch per timed step, counted by the library
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] * 1e9 * r["kernel_ms"] * 1e-3 - r["algorithmic_bytes_per_launch"]) < 1e-3 * r["algorithmic_bytes_per_launch"]
    assert r["traffic"] is None or 0.8 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.2
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert 0.5 < e["frac_of_pcie_ceiling"] < 1.1 and e["engine_scratch_bytes"] < 2 * 2 ** 30
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["batch"]["clips"] == 128 and len(d["batch"]["clip_checksums_sha16"]) == 16
    assert {"value_dense", "value_broadband", "pipeline_u8_dense", "pipeline_u8_dense_fast", "pipeline_u8_broadband"} <= set(d["worst_case"])
    sm = d["scatter_modes"]
    assert {f"{a}_{b}" for a in ("sparse", "dense") for b in ("u64_reds", "f32_reds", "sorted")} <= set(sm)
    assert all(sm[k]["value"] > 0 and sm[k]["scatter_ms"] > 0 for k in sm if isinstance(sm[k], dict))
    assert len(d["nfft_sweep"]) == 8 and d["stream_latency"]["p50_us"] < 1000.0     # the < 1 ms target of north_star
