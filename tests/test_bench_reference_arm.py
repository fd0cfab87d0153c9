"""`bench.py --impl reference` needs no GPU: the CPU arm of the driver's ratio (the C++/OpenMP restatement of the reference's
update(), oracle/cpu_restated.cpp) must print one JSON line with the bench contract's keys; under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "3", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))


def test_reference_arm_line():
    r = run()
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Navier2D timesteps/sec" and d["unit"] == "steps/s"
    assert d["steps"] == 3 and d["warmup"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["config"]["config"] == "C1" and "129x129" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and str(cb["cores"]) in cb["threads_tried_s_per_step"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ops = d["ops"]   # ms / transform, ms / solve of the CPU arm (the metric's second half)
    assert "error" not in ops and all(v > 0 for v in ops["ms_per_transform"].values()) and all(v > 0 for v in ops["ms_per_solve"].values())


def test_reference_arm_other_ranks_do_nothing():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == "", r.stdout + r.stderr[-2000:]
