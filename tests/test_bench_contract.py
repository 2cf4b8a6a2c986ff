"""bench.py output contract on a CPU-only box: the reference arm (`--impl reference`, the oracle port of the reference's CPU
path) prints ONE JSON line with the keys the driver reads; the b200 arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    r = run("--impl", "reference", "--tiny", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(line) == 1
    j = json.loads(line[0])
    assert j["impl"] == "reference" and j["unit"] == "crops/s" and j["higher_is_better"] is True and j["n_gpus"] == 1
    assert j["steps"] == 1 and j["warmup"] == 1 and j["value"] > 0 and j["ms_per_step"] > 0
    assert set(j["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and isinstance(cb["sample"], str)
    assert "workload" in j["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_b200_arm_needs_a_gpu():
    r = run("--tiny", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0                       # fails loudly: there is no CPU / PyTorch fallback for the product path
    assert "{\"metric\"" not in r.stdout
