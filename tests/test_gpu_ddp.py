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
    # own process group + kill of the whole group on a timeout: a hung rank must not survive the test and keep a
    # spinning NCCL kernel on the GPU under whatever runs next
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("ddp_worker timed out (killed):\n" + out[-2000:] + err[-2000:])
    print(out[-3000:], err[-3000:])
    assert proc.returncode == 0 and "[ddp_check] PASS" in out
