"""bench.py's contract on a CPU-only machine: the reference arm prints ONE JSON line with the agreed keys (it times the
CPU oracle - the only place outside tests/ and smoke() that may execute oracle/), and our arm refuses to run without a
GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "graph_vit_train_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "batch" in cb["sample"]
    assert "workload" in d["config"]


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return                                   # on the GPU box the arm runs; covered by the driver itself
    r = _run("--steps", "1", "--warmup", "1", timeout=120)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
