"""CPU: the bench lines kept under profiles/ carry every key of the measurement contract (they are what `bench.py` printed on
the GPU box; this guards the format the driver parses, not the numbers)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not kept")
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def test_default_line_has_the_contract_keys():
    d = _line("r2_bench_large_sphere.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["metric"] == "mesh_extraction_vertices_per_s" and d["unit"] == "vertices/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "large" in d["config"]["workload"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["units"] > 0
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]            # host copies are inside the end-to-end region
    assert d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    assert abs(d["value"] - d["config"]["mesh_vertices"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line("r2_bench_large_sphere_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "mesh_extraction_vertices_per_s" and d["unit"] == "vertices/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    ours = _line("r2_bench_large_sphere.json")
    assert d["config"] == ours["config"]      # both arms describe the same workload with the same keys


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines(n):
    d = _line(f"r2_bench_n{n}_large_sphere.json")
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["run"]["objects_per_step"] == n
    if "rank_ms_mean_min_max" in d["run"]:   # lines taken after the key was added
        assert len(d["run"]["rank_ms_mean_min_max"]) == n
