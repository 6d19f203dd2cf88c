"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints ONE JSON line with the contract's
keys; under a torchrun-style environment only rank 0 prints.  (The GPU arm's line is checked by the driver on the box.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "transitions/s" and d["vs_baseline"] is None
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert cb["replay_only"]["cores"] == 1 and cb["replay_only"]["sample_transitions_per_s"] > 0


def test_reference_arm_other_ranks_stay_silent():
    assert run(dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")) == []
