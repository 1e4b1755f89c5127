"""bench.py contract on CPU: the reference arm prints ONE JSON line with the agreed keys (the CUDA arm needs a
GPU and is exercised by the driver), non-zero ranks of the reference arm stay silent, and the CUDA arm refuses to run
without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_json_line():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--points", "100000", "--hyp", "64"])
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d
    assert d["config"]["workload"].startswith("C1 synthetic curved tunnel") and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "parity unpinned" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_thread_count_ignores_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use (and report) every usable core."""
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--points", "60000", "--hyp", "64"], env={"OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, res.stderr
    d = json.loads(res.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["config"]["points_per_scan"] == 60000 and d["config"]["plane_hypotheses"] == 32


def test_reference_arm_other_ranks_print_nothing():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--points", "50000"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = _run(["--steps", "1", "--warmup", "1"])
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)
