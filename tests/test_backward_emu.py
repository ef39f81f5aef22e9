"""Backward pass of the CUDA path, executed on the CPU through the emulation shim (csrc/emu.h), against
autograd of the oracle (fp64).  These are the same .cu sources the GPU library is built from; the GPU
tests (test_gpu_train.py) repeat the comparison on the device."""
import pytest
import torch

import tdanet_b200.look2hear as look2hear
from oracle import tdanet_oracle as O
import emu_harness as H

SR = 8000


CLASS = {"best": "TDANetBest", "fork": "TDANet", "origin": "TDANetOrigin", "multres": "TDANetMultRes"}


def _model_sd(kw, seed=0, variant="best"):
    torch.manual_seed(seed)
    model = getattr(look2hear.models, CLASS[variant])(sample_rate=SR, **kw)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    # move the parameters that initialise to constants (GlobLN gamma = 1, beta = 0, biases) off their defaults
    for k, v in sd.items():
        if k.endswith("pos_enc.pe"):
            continue
        if k.endswith((".gamma", ".beta", "norm.weight", "norm.bias", "attn_in_norm.weight", "attn_in_norm.bias", ".bias")):
            sd[k] = v + 0.2 * torch.randn(v.shape, generator=g)
    return sd


def _autograd(sd, wav, d_est, kw, variant="best", **drop):
    sd64 = {k: v.double().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    est = O.forward(sd64, wav.double(), O.OracleConfig(variant=variant, sample_rate=SR, **kw, **drop))
    (est * d_est.double()).sum().backward()
    return {k: v.grad for k, v in sd64.items() if not k.endswith("pos_enc.pe")}


CASES = {
    "depth4": dict(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=4, enc_kernel_size=4, num_sources=2),
    "depth5_odd": dict(out_channels=16, in_channels=64, num_blocks=2, upsampling_depth=5, enc_kernel_size=2, num_sources=2),
    "depth2_3src": dict(out_channels=16, in_channels=32, num_blocks=3, upsampling_depth=2, enc_kernel_size=4, num_sources=3),
    "depth3": dict(out_channels=32, in_channels=32, num_blocks=1, upsampling_depth=3, enc_kernel_size=4, num_sources=2),
    # TDANetMultRes: `kernels` encoder convs of window (k+1)*K, no bottleneck, attention over the time axis
    "multres4": dict(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=4, enc_kernel_size=4, num_sources=2, kernels=4),
    "multres2": dict(out_channels=32, in_channels=64, num_blocks=2, upsampling_depth=3, enc_kernel_size=2, num_sources=2, kernels=2),
}


# (lengths kept short: the emulation runs the warp-shuffle kernels - attention backward, LayerNorm rows - as real OS
# threads with barriers, which is what the time of this file goes into; the GPU tests repeat every case at 2x the length)
@pytest.mark.parametrize("variant,name,B,T", [
    ("best", "depth4", 2, 643), ("best", "depth5_odd", 3, 587), ("best", "depth2_3src", 2, 331), ("best", "depth3", 1, 517),
    ("fork", "depth4", 2, 643), ("fork", "depth5_odd", 2, 587), ("fork", "depth2_3src", 2, 331),
    ("origin", "depth4", 2, 643), ("multres", "multres4", 2, 643), ("multres", "multres2", 3, 587)])
def test_emulated_backward_matches_autograd(variant, name, B, T):
    kw = CASES[name]
    sd = _model_sd(kw, variant=variant)
    g = torch.Generator().manual_seed(7)
    wav = torch.randn(B, 1, T, generator=g) * 0.1
    d_est = torch.randn(B, kw["num_sources"], T, generator=g)
    grads, _, _ = H.emu_backward(sd, wav, d_est, kw, SR, variant)
    ref = _autograd(sd, wav, d_est, kw, variant)
    dead = f"loc_glo_fus.{kw['upsampling_depth'] - 1}."
    worst = 0.0
    for k, r in ref.items():
        if r is None:
            # parameters the forward never uses (dead loc_glo_fus of the last scale unless it is the first
            # top-down step's partner): the CUDA path must leave their gradient at zero
            assert dead in k or (kw["num_blocks"] == 1 and "concat_block" in k), k
            assert grads[k].abs().max().item() == 0.0, k
            continue
        scale = r.abs().max().item()
        err = (grads[k].double() - r).abs().max().item()
        rel = err / max(scale, 1e-12)
        worst = max(worst, rel)
        assert rel < 2e-4, f"{k}: max-rel {rel:.3e} (|ref|max {scale:.3e})"
    print(f"{variant}/{name}: worst max-rel gradient error {worst:.2e}")


@pytest.mark.parametrize("variant,name,B,T,dropout,drop_path", [
    ("best", "depth4", 3, 643, 0.1, 0.1), ("best", "depth3", 2, 517, 0.3, 0.0), ("best", "depth4", 4, 643, 0.0, 0.4),
    ("fork", "depth4", 3, 643, 0.2, 0.3), ("origin", "depth4", 2, 643, 0.1, 0.1),
    ("multres", "multres4", 2, 643, 0.2, 0.3), ("multres", "multres2", 3, 587, 0.1, 0.0)])
def test_emulated_backward_with_dropout_masks(variant, name, B, T, dropout, drop_path):
    """Train-mode stochastic layers: with the SAME keep-masks in the workspace and in the oracle,
    the emulated backward pass matches autograd of the oracle through nn.Dropout / attention-weight dropout / DropPath."""
    kw = CASES[name]
    sd = _model_sd(kw, variant=variant)
    g = torch.Generator().manual_seed(9)
    wav = torch.randn(B, 1, T, generator=g) * 0.1
    d_est = torch.randn(B, kw["num_sources"], T, generator=g)
    eng = H.make_engine(kw, SR, variant=variant)
    Lb = eng.latent_lengths(T)[0][-1]
    masks = H.random_drop_masks(B, Lb, kw["in_channels"], 8, kw["num_blocks"], dropout, drop_path,
                                time_axis=variant == "multres")
    if drop_path > 0:
        masks[0]["dp"][0, 0] = 0          # at least one dropped and one kept path
        masks[0]["dp"][1, -1] = 0
        masks[-1]["dp"][0, -1] = 1
    grads, est, _ = H.emu_backward(sd, wav, d_est, kw, SR, variant, dropout, drop_path, masks)
    drop = dict(drop_masks=masks, dropout=dropout, drop_path=drop_path)
    ref = _autograd(sd, wav, d_est, kw, variant, **drop)
    # the masks matter: the deterministic forward gives another output
    est0 = O.forward(sd, wav, O.OracleConfig(variant=variant, sample_rate=SR, **kw))
    assert (est - est0).abs().max().item() > 1e-4 * est0.abs().max().item()
    worst = 0.0
    for k, r in ref.items():
        if r is None:
            assert grads[k].abs().max().item() == 0.0, k
            continue
        scale = r.abs().max().item()
        rel = (grads[k].double() - r).abs().max().item() / max(scale, 1e-12)
        worst = max(worst, rel)
        assert rel < 2e-4, f"{k}: max-rel {rel:.3e} (|ref|max {scale:.3e})"
    print(f"{variant}/{name} dropout {dropout} drop_path {drop_path}: worst max-rel gradient error {worst:.2e}")


def test_emulated_warp_attention_backward(monkeypatch):
    """The one-warp-per-problem attention backward (n <= 16 tokens; what the training batches use on the GPU) under
    emulation: real OS threads with barriers, so it gets one small case of its own (the other cases emulate the
    scratch-buffer kernels, which the GPU takes for n > 16)."""
    import subprocess, sys, os
    code = ("import sys, torch; sys.path.insert(0, 'tests'); import emu_harness as H;"
            "from test_backward_emu import _model_sd, _autograd, CASES, SR;"
            "kw = CASES['depth4']; sd = _model_sd(kw); g = torch.Generator().manual_seed(3);"
            "wav = torch.randn(3, 1, 331, generator=g) * 0.1; d = torch.randn(3, 2, 331, generator=g);"
            "masks = H.random_drop_masks(3, H.make_engine(kw, SR).latent_lengths(331)[0][-1], kw['in_channels'], 8, kw['num_blocks'], 0.2, 0.0);"
            "grads, _, _ = H.emu_backward(sd, wav, d, kw, SR, 'best', 0.2, 0.0, masks);"
            "ref = _autograd(sd, wav, d, kw, 'best', drop_masks=masks, dropout=0.2, drop_path=0.0);"
            "worst = max((grads[k].double() - r).abs().max().item() / max(r.abs().max().item(), 1e-12) for k, r in ref.items() if r is not None);"
            "print('worst', worst); assert worst < 2e-4")
    env = dict(os.environ, TD_EMU_WARP_ATT="1")   # read once per process by the emulation library: fresh interpreter
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("variant", ["best", "fork"])
def test_emulated_backward_with_bf16_stored_activations(variant):
    """act_dtype bf16 in training: the large activations a forward keeps (proj, spp_dw outputs, x_fused, expanded) are
    stored as bf16 and the backward kernels read them through run-time typed loads.  The workspace is filled from the
    oracle's taps rounded to bf16, so the gradients are those of the exact model evaluated at activations that are off
    by up to 2^-9 relative: close to autograd (a loose bound), different from the fp32-storage run (the path is live),
    and the workspace reports 2-byte elements for exactly those tensors."""
    kw = dict(out_channels=32, in_channels=64, num_blocks=2, upsampling_depth=4, enc_kernel_size=4, num_sources=2)
    sd = _model_sd(kw, variant=variant)
    g = torch.Generator().manual_seed(7)
    B, T = 2, 643
    wav = torch.randn(B, 1, T, generator=g) * 0.1
    d_est = torch.randn(B, 2, T, generator=g)
    g16, _, ws16 = H.emu_backward(sd, wav, d_est, kw, SR, variant, act_dtype="bf16")
    g32, _, ws32 = H.emu_backward(sd, wav, d_est, kw, SR, variant)
    import numpy as np
    for name in ("proj", "spp0", "spp3", "fused0", "expanded1"):
        assert ws16.view(name).dtype == np.uint16 and ws32.view(name).dtype == np.float32, name
    for name in ("x0", "y", "ga_out", "fc1", "qkv"):
        assert ws16.view(name).dtype == np.float32, name
    ref = _autograd(sd, wav, d_est, kw, variant)
    num = den = diff = 0.0
    for k, r in ref.items():
        if r is None:
            assert g16[k].abs().max().item() == 0.0, k
            continue
        num += (g16[k].double() - r).pow(2).sum().item()
        diff += (g16[k].double() - g32[k].double()).pow(2).sum().item()
        den += r.pow(2).sum().item()
        assert torch.isfinite(g16[k]).all(), k
    rel, rel_vs_fp32 = (num / den) ** 0.5, (diff / den) ** 0.5
    print(f"{variant} bf16 storage: whole-gradient rel-L2 vs fp64 autograd {rel:.2e}, vs the fp32-storage run {rel_vs_fp32:.2e}")
    assert rel < 3e-2, rel
    assert rel_vs_fp32 > 1e-5, rel_vs_fp32
