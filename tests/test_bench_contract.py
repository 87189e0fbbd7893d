"""The committed bench lines (profiles/r01_bench_*.json, produced by bench.py on a B200) carry every key the
measurement contract asks for.  Guards the JSON shape; the numbers themselves are measured, not tested."""
import json
import os

import pytest

PROFILES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
BASE = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float,
        "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict,
        "gpu_launches": int}


def load(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["r01_bench_n1.json", "r01_bench_n2.json", "r01_bench_n8.json", "r01_bench_n2_strong.json"])
def test_cuda_arm_line(name):
    d = load(name)
    for k, ty in BASE.items():
        assert isinstance(d[k], ty), k
    assert "vs_baseline" in d and d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric
    assert d["unit"] == "Mvoxels/s" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["higher_is_better"]
    assert "prospero" in d["config"]["workload"] and "4096" in d["config"]["workload"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] > 0
    assert d["value"] == pytest.approx(d["config"].get("frames", 1) * (d["n_gpus"] if d["scaling"] == "weak" else 1)
                                       * 4096 * 4096 / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)


def test_n1_line_has_the_cpu_baseline():
    d = load("r01_bench_n1.json")
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["cores"] >= 1
    assert d["n_gpus"] == 1 and d["scaling"] == "weak"


def test_reference_arm_line():
    d = load("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "Mvoxels/s" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
