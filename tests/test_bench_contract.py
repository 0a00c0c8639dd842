"""bench.py's output contract, on the CPU: the reference arm (`--impl reference`) runs here (it times the unmodified reference
vendored in oracle/_ref, or the oracle port) and must print ONE JSON line with the keys the driver reads; the GPU arm must
fail loudly without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                      # the reference's own prints must not reach stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "plane_batch" not in d["config"]      # same config dict as the GPU arm's
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "models", "drmvsnet.py")):
        assert d["cpu_baseline"]["kind"] == "reference"


def test_training_reference_arm():
    r = _run("--impl", "reference", "--workload", "cfg5")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["value"] > 0 and d["config"]["batch"] == 4


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _run("--workload", "tiny", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
