"""Multi-GPU parity (NCCL): runs tests/multigpu_check.py under torchrun when the box has at
least two GPUs (`gpurun --gpus 2`); on a one-GPU box only the one-rank degenerate case runs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_row_sharded_lasso_and_sharded_starts(gpu):
    nproc = 2 if gpu >= 2 else 1
    r = _run(nproc, 29511)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MULTIGPU_OK world={nproc}" in r.stdout
