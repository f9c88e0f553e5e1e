"""The bench line the driver parses: keys of the JSON contract, checked on the committed line of the last
GPU run (profiles/) and on the static pieces of bench.py that do not need a GPU."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def latest_bench_line():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_v*.json")),
                   key=lambda f: (len(os.path.basename(f)), os.path.basename(f)))
    single = [f for f in files if "_n" not in os.path.basename(f)]
    assert single, "no committed bench line under profiles/"
    return json.load(open(single[-1])), single[-1]


def test_committed_bench_line_follows_the_contract():
    line, path = latest_bench_line()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, (path, key)
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["gpu_launches"] >= line["steps"] > 0
    e2e = line["e2e"]
    assert e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert 0.5 * line["value"] <= e2e["value"] <= 1.01 * line["value"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    # achieved = algorithmic flop per launch / the kernel's launch duration
    assert abs(roof["achieved"] - roof["evals_per_launch"] * roof["flop_per_eval"] / (roof["launch_ms"] * 1e-3) / 1e12) < 1e-6
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["sample"]
    clocks = line["clocks"]
    assert clocks["sm_mhz"] > 0.9 * clocks["sm_max_mhz"]
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_traffic_file_matches_the_committed_profile():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    line, _ = latest_bench_line()
    assert line["roofline"]["traffic"] is None or line["roofline"]["traffic"] > 0
    assert t["objective_kernel_dram_bytes_per_launch"] > 0 and 15 < t["objective_kernel_fp64_inst_per_eval"] < 61
    src = t["source"].split(":")[0]
    assert os.path.exists(os.path.join(ROOT, src)), src
