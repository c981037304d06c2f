"""CPU: bench.py's command-line contract where it can be exercised without a GPU — the reference arm
(`--impl reference`) prints exactly one JSON line with the required keys on a tiny workload, and the own
arm refuses to run without CUDA instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=str(ROOT))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "2", "--res", "16", "--params", "5000")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["steps"] == 2


def test_reference_arm_nonzero_rank_exits_quietly(monkeypatch):
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--batch", "2", "--res", "16", "--params", "5000"],
                         capture_output=True, text=True, timeout=300, cwd=str(ROOT), env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_own_arm_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        return
    out = _run("--steps", "1", "--warmup", "0")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
