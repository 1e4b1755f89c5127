"""The segment-form radix sort (csrc/gm_sort.cuh) alone against std::stable_sort: tools/microbench/sort_v2.cu in its
check-only mode -- sizes 1 .. 10M (ragged, bucket-padded grids), key widths 3 .. 32 bits (digit widths 4 .. 8), inputs with
many equal keys (stability).  Replaces `std::sort` of (key, index) pairs in pcl::VoxelGrid
(/root/reference src/tunnel_processing.cpp:215-220); everything downstream of it is covered by the chain parity tests."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_radix_sort_kernels_against_stable_sort(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "sort_v2")
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    mb = os.path.join(ROOT, "tools", "microbench")
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-std=c++17",
                          "-I", os.path.join(ROOT, "geometric_mapping_b200", "csrc"), "-I", mb, "-o", exe, os.path.join(mb, "sort_v2.cu")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    run = subprocess.run([exe, "0"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "total failing cases: 0" in run.stdout, run.stdout[-3000:] + run.stderr[-1000:]
    assert run.stdout.count("mismatches=0 (no error)") >= 50
