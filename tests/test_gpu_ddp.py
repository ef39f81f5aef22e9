"""The batch-sharded training step on 2 GPUs over NCCL (BASELINE.json configs[3]; the reference's only collective
is Lightning DDP's gradient mean, look2hear/system/audio_litmodule.py:83-124 + audio_train.py:187-197).
Runs tests/ddp_worker.py under torch.distributed.run; skipped on a box with a single GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_training_step_matches_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "ddp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "[ddp_check] PASS" in r.stdout
