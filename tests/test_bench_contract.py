"""bench.py's reference arm (`--impl reference`: the CPU port of the reference algorithm on the host cores) prints the
JSON line the driver expects.  Runs on CPU; the GPU arm shares `workload_config`, METRIC and UNIT with it."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "12", "--dirs", "16", "--levels", "2",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cell_direction_updates_per_s" and line["unit"] == "updates/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["steps"] == 2 and line["warmup"] == 3 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["cells"] == 12 ** 3 and line["config"]["directions"] == 16 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "12", "--dirs", "16",
                          "--steps", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
