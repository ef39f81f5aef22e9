"""2+-rank check of the batch-sharded training step (run under torch.distributed.run, one rank per GPU, by
tests/test_gpu_ddp.py): every rank builds its model from a DIFFERENT seed - TrainingStep must broadcast rank 0's
parameters like DistributedDataParallel's constructor does in the reference run (audio_train.py:187-197 never
seeds) - then the all-reduced gradient equals the single-process gradient of the concatenated batch (attention
grouped per shard, like reference DDP where every rank attends within its own batch), parameters stay identical
across ranks after optimiser steps, and the loss goes down."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import torch.distributed as dist

import tdanet_b200.look2hear as look2hear


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    kw = dict(out_channels=32, in_channels=64, num_blocks=3, upsampling_depth=4, enc_kernel_size=4, num_sources=2)
    B, T = 4, 4000
    L = look2hear.losses

    def make(seed):
        torch.manual_seed(seed)
        m = look2hear.models.TDANetBest(sample_rate=8000, **kw).to(dev).train()
        m.gemm_mode = "fp32"
        m.dropout = m.drop_path = 0.0    # the deterministic step: 1-GPU and 2-GPU gradients are comparable
        return m

    g = torch.Generator().manual_seed(123)
    tgt_all = torch.randn(world * B, 2, T, generator=g) * 0.1
    mix_all = tgt_all.sum(1)
    tgt, mix = tgt_all[rank * B:(rank + 1) * B].to(dev), mix_all[rank * B:(rank + 1) * B].to(dev)
    loss_fn = L.PITLossWrapper(L.pairwise_neg_sisdr, threshold_byloss=True)
    ts = look2hear.system.TrainingStep(make(1000 + rank), loss_fn, lr=1e-3, clip_grad_norm=5.0)
    # the constructor broadcast rank 0's parameters: every rank now holds the seed-1000 model
    p0 = ts.params.flat.clone()
    dist.broadcast(p0, src=0)
    started_equal = torch.tensor([int(bool((p0 == ts.params.flat).all().item()))], device=dev)
    dist.all_reduce(started_equal, op=dist.ReduceOp.MIN)
    # ---- gradient of one step against the single-process run on the whole batch
    ts.params.zero_grad()
    loss = ts.forward_backward(mix, tgt)
    dist.all_reduce(ts.params.grad, op=dist.ReduceOp.SUM)
    grad_ddp = ts.params.grad / world
    loss_mean = loss.clone()
    dist.all_reduce(loss_mean, op=dist.ReduceOp.SUM)
    loss_mean /= world
    ok = True
    solo = dist.new_group(ranks=[0])           # (collective: every rank calls it) a one-rank group for the reference run
    if rank == 0:
        m1 = make(1000)
        m1.attn_group = B                      # every shard attends within itself
        # a single-process TrainingStep beside the distributed one: its group has one member, so its constructor
        # broadcast and its gradient all-reduce are no-ops (with the default group rank 0 would wait for rank 1 forever)
        ts1 = look2hear.system.TrainingStep(m1, loss_fn, lr=1e-3, clip_grad_norm=5.0, process_group=solo)
        ts1.params.zero_grad()
        loss1 = ts1.forward_backward(mix_all.to(dev), tgt_all.to(dev))
        rel = ((grad_ddp - ts1.params.grad).norm() / ts1.params.grad.norm()).item()
        dl = abs(loss1.item() - loss_mean.item())
        print(f"[ddp_check] world {world}: all-reduced gradient vs single process rel-L2 {rel:.2e}, loss diff {dl:.2e}")
        print(f"[ddp_check] ranks seeded differently start from rank 0's parameters: {bool(started_equal.item())}")
        ok = ok and rel < 1e-4 and dl < 1e-4 and bool(started_equal.item())
    # ---- a few optimiser steps: replicas stay bit-identical, loss decreases
    losses = []
    for _ in range(10):
        losses.append(ts.step(mix, tgt))
    lmean = torch.stack([l.reshape(()) for l in losses])
    dist.all_reduce(lmean, op=dist.ReduceOp.SUM)
    lmean /= world
    ref = ts.params.flat.clone()
    dist.broadcast(ref, src=0)
    same = bool((ref == ts.params.flat).all().item())
    flags = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[ddp_check] replicas identical after 10 steps: {bool(flags.item())}; mean loss {lmean[0].item():.4f} -> {lmean[-1].item():.4f}")
        ok = ok and bool(flags.item()) and lmean[-1].item() < lmean[0].item()
        print("[ddp_check] PASS" if ok else "[ddp_check] FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
