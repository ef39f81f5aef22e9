"""Diagnostic: where does the TF32-mode gradient error come from?  (forward mode x backward mode) grid."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import tdanet_b200.look2hear as look2hear
from test_backward_emu import CASES, SR, _autograd, _model_sd
from test_gpu_train import _model, _inputs, DEV

for name, B, T in [("depth5_odd", 3, 1111), ("depth3", 1, 997), ("depth4", 2, 1203)]:
    kw = CASES[name]
    sd = _model_sd(kw)
    wav, d_est = _inputs(kw, B, T)
    ref = _autograd(sd, wav, d_est, kw)
    for fmode in ("fp32", "tf32"):
        for bmode in ("fp32", "tf32"):
            m = _model(kw, sd).train()
            eng = m.engine
            named = [(n, p) for n, p in m.named_parameters()]
            m.gemm_mode = fmode
            x = wav.squeeze(1).to(DEV)
            est = eng.forward_train(m._weights(), x)
            views = {n: torch.zeros_like(p) for n, p in named}
            gw = eng.pack(views, optional=True)
            m.gemm_mode = bmode
            eng.backward(m._weights(), gw, x, d_est.to(DEV))
            torch.cuda.synchronize()
            errs = []
            num = den = 0.0
            for n, _ in named:
                r = ref[n]
                if r is None:
                    continue
                d = views[n].cpu().double() - r
                errs.append((d.abs().max().item() / max(r.abs().max().item(), 1e-12), n, r.abs().max().item()))
                num += d.pow(2).sum().item(); den += r.pow(2).sum().item()
            errs.sort(reverse=True)
            print(f"{name} fwd={fmode} bwd={bmode}: whole rel-L2 {(num/den)**0.5:.2e}; worst: " +
                  ", ".join(f"{n.replace('sm.unet.','')}={e:.1e}(|g|{s:.1e})" for e, n, s in errs[:4]))
