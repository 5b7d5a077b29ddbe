"""Multi-GPU correctness over NCCL with the real kernels: sharded == full batch for the losses,
the centres, Sinkhorn-Knopp, and DDP-averaged head gradients.  Launches tools/dist_parity_nccl.py under
torch.distributed.run on 2 GPUs (and on every GPU of the box when there are more); skipped on a single-GPU box.
The output is kept under gpurun_out/ so a run on a multi-GPU box leaves its evidence behind."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + n), os.path.join(ROOT, "tools", "dist_parity_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dist_parity_{n}gpu.txt"), "w") as f:
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    return r


@pytest.mark.parametrize("n", [2, 8])
def test_sharded_equals_full_batch_over_nccl(n):
    have = torch.cuda.device_count()
    if have < n:
        pytest.skip(f"needs {n} GPUs, this box has {have}")
    r = _run(n)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "DIST PARITY PASS" in r.stdout, r.stdout[-3000:]
    assert "FAIL" not in r.stdout
