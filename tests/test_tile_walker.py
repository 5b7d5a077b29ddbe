"""Tile schedules of the persistent GEMM kernels, checked on the host: `TileWalker` (csrc/gemm_core.cuh) is
`__host__ __device__`, so a small host program enumerates what every cluster walks for thousands of
(M tiles, N tiles, batches / splits, cluster count) combinations and all three walks (strided, contiguous
runs for the resident-A mode, columns for the resident-B mode): every tile exactly once, balanced load."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


@pytest.mark.skipif(_nvcc() is None, reason="nvcc not found")
def test_every_tile_is_walked_exactly_once(tmp_path):
    exe = str(tmp_path / "walker_check")
    cmd = [_nvcc(), "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "dinox_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "native", "walker_check.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
