"""bench.py contract on a CPU box: the reference arm prints exactly one JSON line on stdout with the keys the driver reads
(the GPU arm needs a B200 and is exercised by the driver itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_port_arm_matches_reference_arm_interface():
    """The GPU box has no /root/reference: the arm must fall back to the oracle port (kind "port") with the same line."""
    env = dict(os.environ, DTQN_REFERENCE_ROOT="/nonexistent")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    # >= 200 timed iterations whatever --steps says, the 1-thread figure, and the three separate legs
    assert "50000 prepopulate" in cb["sample"] and int(cb["sample"].split()[0]) >= 200 and cb["one_thread_value"] > 0
    assert set(cb["legs"]) == {"prepopulate_env_steps_per_sec_1thread", "run_step_env_steps_per_sec",
                               "agent_train_grad_steps_per_sec", "loop_iterations_per_sec"}
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
