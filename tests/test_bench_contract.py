"""bench.py's reference arm runs without a GPU: its stdout must be exactly one JSON line with the contract's keys, even
when something in the process writes to file descriptor 1 on its own (as NCCL does with its version line)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    env.pop("RANK", None)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    # the arm reports what it executed: `steps` real sample passes whose duration is ms_per_step, the extrapolation flagged
    assert line["steps"] == 1 and line["extrapolated"] is False and line["sample_factor"] == 1.0   # 20 000 rows < the sample size
    assert line["ms_per_step"] == line["cpu_baseline"]["sample_seconds_per_step"] * 1e3
    assert line["anchor"]["extrapolated"] is False and line["anchor"]["rows"] == 20000


def test_reference_arm_flags_its_extrapolation():
    env = dict(os.environ)
    env.pop("RANK", None)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "40000", "--cpu-sample-rows",
                           "10000", "--steps", "2", "--warmup", "1", "--no-anchor"], capture_output=True, text=True, timeout=600,
                          env=env, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip())
    assert line["steps"] == 2 and line["warmup"] == 1 and line["extrapolated"] is True and line["sample_factor"] == 0.25
    assert "anchor" not in line


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--rows", "20000",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert proc.returncode == 0 and proc.stdout.strip() == ""
