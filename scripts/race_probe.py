"""Repeats the graph-replayed B=64 forward and reports the spread of its distance to the fp32-GEMM result:
a data race between kernels shows up as replays that differ by more than the TF32 noise (a few 1e-4)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import tdanet_b200.look2hear as look2hear
from bench import model_kwargs, SR

dev = "cuda:0"
variant = sys.argv[1] if len(sys.argv) > 1 else "best"
enc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
m = getattr(look2hear.models, {"best": "TDANetBest", "fork": "TDANet"}[variant])(sample_rate=SR, **model_kwargs(enc)).eval().to(dev)
x = (torch.randn(64, 1, 32000, generator=torch.Generator().manual_seed(1234)) * 0.1).to(dev)
with torch.no_grad():
    m.gemm_mode = "fp32"
    ref = m(x).clone()
    m.gemm_mode = "tf32"
    scale = ref.abs().max().item()
    errs = []
    for i in range(3):
        errs.append(((m(x) - ref).abs().max().item() / scale))
    m.use_cuda_graph = True
    gerrs = []
    for i in range(12):
        gerrs.append(((m(x) - ref).abs().max().item() / scale))
print(f"{variant} {enc} ms  env TMA_STORE={os.environ.get('TDANET_GEMM_TMA_STORE','1')} PDL={os.environ.get('TDANET_PDL','1')}: "
      f"eager {['%.1e' % e for e in errs]}  graph max {max(gerrs):.1e} min {min(gerrs):.1e}")
