"""The multi-process, multi-GPU parity check as a pytest test: when the box has >= 2 GPUs it launches one rank per GPU
(torch.distributed.run, NCCL bootstrap, peer exchange over NVLink) running tests/multi_gpu_check.py, which compares the
row-sharded Ell with the CPU checker — bit for bit in reduce_mode 0, 1e-9 in the blocked modes, including the n = 16384 /
reduce_mode = world / p2p configuration bench.py times — and the problem-sharded batch with the single-rank result."""
import os
import socket
import subprocess
import sys

import pytest

from ellalgo_rs_b200._lib import lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_ranks_against_checker():
    nd = lib().ellcuda_device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if nd >= 8 else 4 if nd >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0 and "SHARDED CHECK PASSED" in p.stdout, tail
