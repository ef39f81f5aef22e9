"""Repeats the pipelined end-to-end path of bench.py and prints every repetition's ms per step (bimodality probe)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import tdanet_b200.look2hear as look2hear
import tdanet_b200.look2hear.system as L2S
from bench import model_kwargs, SR

dev = torch.device("cuda", 0)
variant = sys.argv[1] if len(sys.argv) > 1 else "best"
act = sys.argv[2] if len(sys.argv) > 2 else "fp32"
torch.manual_seed(0)
m = getattr(look2hear.models, {"best": "TDANetBest", "fork": "TDANet"}[variant])(sample_rate=SR, **model_kwargs(4)).eval().to(dev)
m.act_dtype = act
m.use_cuda_graph = True
B, K = 64, 20
x_host = (torch.randn(B, 1, 32000, generator=torch.Generator().manual_seed(1234)) * 0.1).pin_memory()
outs = [torch.empty(B, 2, 32000).pin_memory() for _ in range(2)]
x_dev = x_host.to(dev)
with torch.no_grad():
    for _ in range(3):
        m(x_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        m(x_dev)
    e1.record(); torch.cuda.synchronize()
    print(f"{variant} {act}: device {e0.elapsed_time(e1) / K:.2f} ms per step")
    L2S.separate_pipelined(m, [x_host] * 3, outs)
    res = []
    for rep in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        L2S.separate_pipelined(m, [x_host] * K, outs)
        e1.record(); e1.synchronize()
        res.append((e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1e3 / K))
    print("pipelined ms per step (events, wall):", [f"{a:.2f}/{b:.2f}" for a, b in res])
