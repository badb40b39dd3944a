"""The bench line contract (task statement ④ + base contract): checked on the committed lines under profiles/ that the
current bench.py wrote, so a field that goes missing is caught on CPU."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_bench_final_*.json")) + glob.glob(os.path.join(ROOT, "profiles", "r1_bench_2gpu_c3.json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_lines_follow_the_contract(path):
    d = json.loads(open(path).read())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["dtype"] == "f32" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) <= 1e-3 * d["value"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert "traffic" in r
    c = d["cpu_baseline"]
    assert c is None or (c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"])
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_bench_defaults_and_reference_arm_are_wired():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'default="c3"' in src and '"--impl"' in src and '"reference"' in src and "args.warmup = max(args.warmup, 3)" in src
