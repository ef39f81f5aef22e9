"""Times the sections of one training step (eager launches, CUDA events, median of 10)."""
import os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import tdanet_b200.look2hear as look2hear
from tdanet_b200.engine import pit_loss
from bench import model_kwargs, train_targets, SR

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = look2hear.models.TDANetBest(sample_rate=SR, **model_kwargs(4)).to(dev).train()
model.gemm_mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
L = look2hear.losses
ts = look2hear.system.TrainingStep(model, L.PITLossWrapper(L.pairwise_neg_sisdr, threshold_byloss=True))
mix, tgt = train_targets(0, 8)
mix, tgt = mix.to(dev), tgt.to(dev)
eng = model.engine
w, gw = ts._pack()
def ev(): return torch.cuda.Event(enable_timing=True)
rows = {k: [] for k in ("zero", "forward", "loss", "backward", "optim", "total")}
for it in range(14):
    e = [ev() for _ in range(6)]
    e[0].record(); ts.params.zero_grad()
    e[1].record(); est = eng.forward_train(w, mix, 0)
    e[2].record(); loss, _, _, d_est = pit_loss(est, tgt, "sisdr", True, want_grad=True)
    e[3].record(); eng.backward(w, gw, mix, d_est, 0)
    e[4].record(); ts.optimizer_step()
    e[5].record(); torch.cuda.synchronize()
    if it >= 4:
        for i, k in enumerate(("zero", "forward", "loss", "backward", "optim")):
            rows[k].append(e[i].elapsed_time(e[i + 1]))
        rows["total"].append(e[0].elapsed_time(e[5]))
print({k: round(statistics.median(v), 3) for k, v in rows.items()})
# inference forward at the same batch for comparison
model.eval()
with torch.no_grad():
    wi = model._weights()
    for _ in range(3): eng.forward(wi, mix)
    a, b = ev(), ev()
    a.record()
    for _ in range(10): eng.forward(wi, mix)
    b.record(); torch.cuda.synchronize()
    print("inference forward B=8 (eager):", round(a.elapsed_time(b) / 10, 3), "ms")
    for _ in range(3): eng.forward_graphed(wi, mix)
    a.record()
    for _ in range(10): eng.forward_graphed(wi, mix)
    b.record(); torch.cuda.synchronize()
    print("inference forward B=8 (graph):", round(a.elapsed_time(b) / 10, 3), "ms")
