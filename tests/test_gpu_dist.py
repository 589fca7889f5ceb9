"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): the NCCL data-parallel step, launched like bench.py is
(torch.distributed.run, one process per GPU).  The checks live in tests/dist_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_gpu_data_parallel_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
