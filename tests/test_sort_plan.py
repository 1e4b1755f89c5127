"""Host-side bookkeeping of the segment-form radix sort (csrc/gm_sort.cuh) without a GPU: tests/cpp/sort_plan_check.cu is
compiled with nvcc (cross-compiles here) and run on the CPU -- it calls only host functions of the header (the plan, the
segment geometry) and replays the position arithmetic of the two kernels against std::stable_sort.  The kernels themselves
are covered by the -m gpu parity tests (every exact output downstream of the sort) and tools/microbench/sort_v2.cu."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sort_plan_and_position_arithmetic(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "sort_plan_check")
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    res = subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "geometric_mapping_b200", "csrc"),
                          "-o", exe, os.path.join(ROOT, "tests", "cpp", "sort_plan_check.cu")], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "ok (0 failures)" in run.stdout, run.stdout[-3000:]
