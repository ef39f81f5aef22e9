"""Diagnostic: which tensor of the training workspace first departs from the oracle (same keep-masks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import emu_harness as H
from oracle import tdanet_oracle as O
from test_backward_emu import CASES, SR, _model_sd
from test_gpu_train import _model, _inputs, _read_masks, _oracle_masks, _CompareWorkspace, DEV

for (B, dropout, drop_path) in [(12, 0.1, 0.0), (12, 0.1, 0.1), (4, 0.1, 0.0), (12, 0.0, 0.0), (12, 0.0, 0.1), (9, 0.1, 0.0), (8, 0.1, 0.0)]:
    kw = CASES["depth4"]
    sd = _model_sd(kw)
    m = _model(kw, sd).train()
    m.dropout, m.drop_path = dropout, drop_path
    m.gemm_mode = "fp32"
    m.manual_seed(42)
    wav, d_est = _inputs(kw, B, 1203)
    m._sync_dropout()
    with torch.no_grad():
        est = m.engine.forward_train(m._weights(), wav.squeeze(1).to(DEV))
    torch.cuda.synchronize()
    masks = _oracle_masks(_read_masks(m, B, 1203, kw["num_blocks"])) if (dropout or drop_path) else None
    cfg = O.OracleConfig(sample_rate=SR, taps={}, tap_all=True, drop_masks=masks, dropout=dropout, drop_path=drop_path, **kw)
    with torch.no_grad():
        ref = O.forward(sd, wav, cfg)
    err = (est.cpu() - ref).abs().max().item() / ref.abs().max().item()
    cmp = _CompareWorkspace(m.engine, B, 1203)
    H.fill_workspace(cmp, cfg.taps, kw, "best")
    bad = [(k, f"{v:.1e}") for k, v in cmp.worst.items() if v > 5e-5]
    print(f"B={B} p={dropout}/{drop_path}: est err {err:.2e}; first bad tensors: {bad[:8]}", flush=True)
