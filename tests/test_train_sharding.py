"""Host-side logic of the batch-sharded training step on CPU (gloo, world_size 2): the flat-gradient sum
all-reduce followed by the 1/world scale the optimiser kernel folds in is the DDP gradient mean, and every
rank ends with the same buffer.  (The kernels themselves are covered by test_backward_emu / test_gpu_train.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tdanet_b200.look2hear.system import shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1003, generator=g)               # this rank's gradient of its shard's mean loss
    mine = flat.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)         # what TrainingStep.optimizer_step does
    scaled = flat * (1.0 / world)                       # grad_scale handed to tdanet_adam_step
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    out[rank] = (scaled, torch.stack(gathered).mean(0))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_is_the_ddp_mean():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        scaled, mean = out[r]
        assert torch.allclose(scaled, mean, atol=1e-7)
    assert torch.equal(out[0][0], out[1][0])


def test_training_shards_cover_the_global_batch():
    # global batch 8*N split into contiguous per-rank shards of 8 (BASELINE.json configs[3])
    for world in (1, 2, 4, 8):
        bounds = [shard_bounds(8 * world, world, r) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 8 * world
        assert all(hi - lo == 8 for lo, hi in bounds)
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))


def _bcast_worker(rank, world, port, out):
    import types
    from tdanet_b200.look2hear.system.training import TrainingStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(500 + rank)        # every rank starts from its own random state
    ts = TrainingStep.__new__(TrainingStep)              # host logic only: no CUDA model behind it
    ts.group = None
    ts.params = types.SimpleNamespace(flat=torch.randn(37, generator=g))
    ts.exp_avg, ts.exp_avg_sq = torch.randn(37, generator=g), torch.rand(37, generator=g)
    ts.step_count = torch.tensor([rank * 10 + 3], dtype=torch.int32)
    ts.broadcast_state()
    out[rank] = (ts.params.flat.clone(), ts.exp_avg.clone(), ts.exp_avg_sq.clone(), int(ts.step_count.item()))
    dist.destroy_process_group()


def test_training_state_is_broadcast_from_rank_zero():
    """What DistributedDataParallel's constructor does for the reference (audio_train.py never seeds): parameters,
    Adam moments and the step counter of rank 0 reach every rank before the first step."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bcast_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(500)
    want = (torch.randn(37, generator=g), torch.randn(37, generator=g), torch.rand(37, generator=g), 3)
    for r in range(world):
        for a, b in zip(out[r][:3], want[:3]):
            assert torch.equal(a, b)
        assert out[r][3] == 3
