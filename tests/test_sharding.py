"""Batch sharding across ranks with world_size-2 gloo processes on CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tdanet_b200.look2hear.system import shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 640):
        for ws in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tdanet_b200.look2hear.system import shard_bounds as sb
    n = 7
    lo, hi = sb(n, world, rank)
    x = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    local = x[lo:hi] * 2          # stands for "separate my shard"
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    parts = [torch.zeros(int(s), 3) for s in sizes]
    # ragged gather by padding to the max shard
    pad = torch.zeros(int(max(sizes)), 3)
    pad[: hi - lo] = local
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    full = torch.cat([o[: int(s)] for o, s in zip(outs, sizes)])
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({"full": full, "tmax": t}, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_gather(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    res = torch.load(out)
    assert torch.equal(res["full"], torch.arange(21, dtype=torch.float32).view(7, 3) * 2)
    assert res["tmax"].item() == 2.0


def test_css_chunking_matches_oracle():
    """Chunk starts / padding of the long-form runner against the oracle restatement of LibriCSSDataset."""
    from oracle import tdanet_oracle as O
    from tdanet_b200.look2hear.system import css_segments
    for n, seg, ov in [(960000, 32000, 0.25), (64000, 32000, 0.25), (70001, 32000, 0.25), (5000, 2000, 0.5), (1999, 2000, 0.25)]:
        starts, pad = css_segments(n, seg, ov)
        segs, pad_o = O.css_segments(torch.arange(n, dtype=torch.float32), seg, ov)
        assert pad == pad_o and len(starts) == segs.shape[0]
        for s, row in zip(starts, segs):
            assert row[0].item() == float(s)
    assert len(css_segments(960000, 32000, 0.25)[0]) == 40      # BASELINE config #5: 40 chunks per 60 s stream
