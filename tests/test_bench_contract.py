"""bench.py really runs: the CPU (reference) arm here, the GPU arm under -m gpu -- small sizes, the full JSON contract."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*args, timeout=900):
    env = dict(os.environ)
    env.pop("RANK", None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                         timeout=timeout, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[-2000:]          # ONE JSON line on stdout, nothing else
    return json.loads(lines[0])


def test_reference_arm_runs_and_keeps_the_contract():
    line = run_bench("--impl", "reference", "--grid", "48", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"].startswith("k-point*candidate objective evals/s") and line["unit"] == "evals/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    assert line["e2e"] == dict(value=line["value"], unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in line["config"] and "model" not in line["config"]
    # both arms describe the workload with the same config dict (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(batch=4096, grid=48, gpus=1)
    assert line["config"] == bench.config_dict(a)


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_runs_and_keeps_the_contract():
    line = run_bench("--batch", "64", "--grid", "64", "--grid5", "96", "--steps", "2", "--warmup", "3", "--quick-extras")
    assert BASE_KEYS | {"roofline", "clocks", "parity", "extra"} <= set(line)
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["gpu_launches"] == 2
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 64 * 13 * 8 and line["e2e"]["d2h_bytes_per_step"] == 4 * 64 * 8
    assert "norm_batch" in line["e2e"]["api"]
    r = line["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 30 < r["peak"] < 39
    p = line["parity"]
    assert p["n"] == 8 and p["max_rel_S"] <= 1e-12 and p["extrema_bitwise"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0
    x = line["extra"]
    assert set(x) == {"config1", "config2", "config3_scan", "config3_optimize", "config5_slabs"}
    assert all("error" not in v for v in x.values()), x
    assert x["config2"]["omega_max_rel_err"] <= 1e-13 and x["config2"]["norm_rel_err"] <= 1e-12
    assert x["config3_scan"]["variants_identical"] and x["config3_scan"]["parity"]["extrema_bitwise"]
    assert x["config5_slabs"]["parity"]["extrema_bitwise"] and x["config5_slabs"]["parity"]["max_rel_S"] <= 1e-12
    assert abs(x["config1"]["norm_minus_reference"]) <= 3e-6
