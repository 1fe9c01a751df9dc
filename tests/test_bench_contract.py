"""bench.py's reference arm (CPU: the oracle port timed on the host cores) prints exactly one JSON line that carries
the contract's keys; under a multi-rank launch only rank 0 prints.  (The CUDA arm is exercised on the GPU box.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, gpus=1):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shape", "toy",
                          "--cpu-budget", "0.5", "--gpus", str(gpus)], capture_output=True, text=True, env=env,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_rows_per_s" and d["unit"] == "train rows/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "eval_value": d["eval"]["value"]}
    assert "workload" in d["config"] and d["data"] == "synthetic" and d["scaling"] == "weak"


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, gpus=2) == []
    lines = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, gpus=2)
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and "B=64" in d["config"]["workload"]       # data-parallel arm: global batch 2 x 32
