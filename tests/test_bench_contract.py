"""The bench.py contract: one JSON line with the agreed keys, for the reference arm (CPU, runs here) and for our
arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "4", "--warmup", "3")
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "env-steps/sec (whole box)" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 1e6 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                                             "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["workload"] and "model" not in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "4", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--steps", "64", "--warmup", "8", "--cpu-seconds", "1")
    assert COMMON <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 64 and d["warmup"] == 8 and d["scaling"] == "weak" and d["dtype"] == "f64"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] * 1e9 - 46 * 2097152 / (d["ms_per_step"] * 1e-3)) / (r["achieved"] * 1e9) < 1e-6
    assert d["gpu_launches"] == 64 and d["value"] > 1e10 and d["timed_regions"] >= 5
    assert r["frac_one_stream"] > 0 and r["frac_isolated"] > 0 and d["timing"]["regions"] == d["timed_regions"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 2097152 and e["d2h_bytes_per_step"] > 4 * 2097152 and 0 < e["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["python_loop_value_per_core"] > 0
    assert {"policy_step", "step_with_observation", "mcts_rollouts", "mcts_choose_action", "config_4096_envs",
            "step_many_2Mi_envs", "dqn_loop"} <= set(d["extras"])
    assert "error" not in d["extras"]["dqn_loop"] and d["e2e"]["roofline"]["peak"] > 0
