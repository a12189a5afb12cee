"""Multi-GPU parity (SURVEY.md §4 'distributed' row) as a pytest: launches tests/dist_nccl_check.py on 2 GPUs of the
box -- dynamic filters under SyncBatchNorm x2 == single-process full batch, host-sync-free SyncBatchNorm == torch's,
the NVLink peer-memory exchange (csrc/ud_comm.cu) == NCCL (eager and CUDA-graph replays), FlatGradients == DDP.
Skipped when the box has fewer than 2 GPUs; the log of the last 2-GPU run is kept under profiles/."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_parity():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_nccl_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": OK") >= 6
