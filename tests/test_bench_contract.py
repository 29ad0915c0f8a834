"""CPU: the parts of bench.py's contract that can be exercised without a GPU — the reference arm's JSON line (the CPU
port of the path timed on the host cores) and the refusal of the product arm to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg1_esm2_t6_llama1b")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "contrastive_step_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "pairs" in cb["sample"]
    from oracle.reference_loader import reference_available
    assert cb["kind"] == ("reference" if reference_available() else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg1_esm2_t6_llama1b" and "model" not in d["config"]


def test_reference_arm_runs_on_rank_zero_only():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2",
               env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_arm_refuses_to_run_without_cuda():
    out = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-e2e")
    assert out.returncode != 0 and "no CPU fallback" in (out.stdout + out.stderr)
