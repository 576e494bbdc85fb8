"""The CPU arm of bench.py (`--impl reference`): runs here without a GPU; its JSON line must carry the contract's keys, report what
was MEASURED as the value (never the extrapolation), and name the sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_measures_config1_in_full():
    d = _run("--workload", "c1", "--steps", "1", "--warmup", "0")
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "ms per full make-tree" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["config"]["reference_sample"] == "whole workload" and d["config"]["reference_sample_cells"] == 5000
    assert d["extrapolated_full_workload_ms"] is None                      # nothing to extrapolate: c1 is run whole
    assert d["value"] == d["ms_per_step"] == d["cpu_baseline"]["value"] == d["e2e"]["value"] and d["value"] > 100.0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_reports_the_sample_not_the_extrapolation():
    d = _run("--workload", "c3", "--steps", "1", "--warmup", "0", "--cpu-sample-cells", "1500")
    assert d["config"]["reference_sample_cells"] == 1500 and "1500 of 1000000 cells" in d["config"]["reference_sample"]
    ex = d["extrapolated_full_workload_ms"]
    assert ex is not None and "NOT a measurement" in ex["how"]
    assert abs(ex["value"] / d["value"] - 1000000 / 1500) < 1e-6           # the labelled extra key, never the value
    # other ranks of a torchrun launch have nothing to do
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=60, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
