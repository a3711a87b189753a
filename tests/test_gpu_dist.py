"""Row-partitioned SG path on two GPUs, one process per GPU: spawns
tests/dist_check_2gpu.py under torch.distributed.run and requires every rank to match the oracle bit for bit.
Skipped where fewer than two GPUs are visible; the single-GPU suite covers the same kernels, peer stores and
residual slots through the single-process group (test_gpu_parity.py::test_sg_row_partitioned_group_bit_exact)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_partitioned_sg_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_check_2gpu.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420)
    tail = (p.stdout + p.stderr)[-4000:]
    assert p.returncode == 0, tail
    assert "MISMATCH" not in p.stdout and "False" not in p.stdout, tail
