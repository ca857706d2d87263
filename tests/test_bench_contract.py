"""The JSON-line contract of bench.py, checked without a GPU: the reference arm (the unmodified reference from
baseline/_ref when it is installed, else the CPU oracle port) really runs here on a tiny sample, and the committed line of the last GPU run (profiles/r2_final_bench.json) carries every key
the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_valid_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "2",
                          "--height", "96", "--width", "128", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()
    lines = [ln for ln in out if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert COMMON <= set(d) and d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "MP/s" and "workload" in d["config"]
    from tools import ref_loader
    expected = "reference" if ref_loader.reference_root() else "port"
    assert d["cpu_baseline"]["kind"] == expected and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_falls_back_to_the_port_without_the_reference(tmp_path):
    env = dict(os.environ, SHINESTACKER_REF="", SS_NO_REFERENCE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "2",
                          "--height", "96", "--width", "128", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=env, check=True).stdout.strip().splitlines()
    d = json.loads([ln for ln in out if ln.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "port" and "absent" in d["cpu_baseline"]["sample"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--frames", "2", "--height", "96", "--width", "128", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=env, check=True).stdout
    assert "{" not in out


def test_committed_gpu_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r2_final_bench.json")
    if not os.path.isfile(path):
        pytest.skip("no committed bench line")
    d = json.loads(open(path).read())
    assert COMMON | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] >= 64 * 4096 * 3072 * 3 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert "workload" in d["config"] and "model" not in d["config"]
