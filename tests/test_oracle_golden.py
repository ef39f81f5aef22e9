"""The oracle (oracle/tdanet_oracle.py) against fixtures generated from the unmodified reference
by oracle/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import CLASSES, golden_state_dict, load_golden, max_rel, oracle_cfg
from oracle import tdanet_oracle as O

TAPS = {  # reference module (forward hook, first call = block 0) -> oracle tap name
    "sm.unet.spp_dw.4": "spp.4",
    "sm.unet.globalatt": "ga.out",
    "sm.unet.last_layer.3": "expanded.3",
    "sm.unet.last_layer.0": "expanded.0",
    "sm.unet": "block.0",
}


@pytest.mark.parametrize("variant", list(CLASSES))
def test_small_model_matches_reference(variant):
    g = load_golden(f"{variant}_small")
    sd = golden_state_dict(g)
    cfg = oracle_cfg(variant, g["kwargs"], g["sample_rate"], taps={})
    with torch.no_grad():
        y = O.forward(sd, torch.from_numpy(g["x"]), cfg)
    assert y.shape == g["y"].shape
    assert max_rel(y, torch.from_numpy(g["y"])) < 2e-6
    for ref_name, tap in TAPS.items():
        ref = torch.from_numpy(g["tap/" + ref_name])
        assert max_rel(cfg.taps[tap], ref) < 2e-6, ref_name


@pytest.mark.parametrize("variant", list(CLASSES))
def test_small_model_input_ranks(variant):
    g = load_golden(f"{variant}_small")
    sd = golden_state_dict(g)
    cfg = oracle_cfg(variant, g["kwargs"], g["sample_rate"])
    x = torch.from_numpy(g["x"])[:1]
    with torch.no_grad():
        y3 = O.forward(sd, x, cfg)
        y2 = O.forward(sd, x[:, 0], cfg)
        y1 = O.forward(sd, x[0, 0], cfg)
    assert y3.shape == (1, 2, x.shape[-1]) and y1.shape == (2, x.shape[-1])
    assert torch.equal(y3, y2) and torch.equal(y3[0], y1)


@pytest.mark.parametrize("variant", list(CLASSES))
def test_full_model_seeded(variant):
    """BASELINE.json configuration, weights re-created from the seed by the product classes (whose
    init is pinned to the reference by test_models_api), output against the stored reference output."""
    import tdanet_b200.look2hear.models as M
    g = load_golden(f"{variant}_full")
    torch.manual_seed(int(g["init_seed"]))
    m = M.get(CLASSES[variant])(sample_rate=g["sample_rate"], **g["kwargs"])
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    B = int(g["batch"])
    x = torch.randn(B, 1, 32000, generator=torch.Generator().manual_seed(int(g["input_seed"]))) * 0.1
    with torch.no_grad():
        y = O.forward(sd, x, oracle_cfg(variant, g["kwargs"], g["sample_rate"]))
    ref = torch.from_numpy(g["y_sub"])
    assert (y[:, :, ::16] - ref).abs().max().item() / float(g["y_absmax"]) < 5e-6


def test_loss_known_answers():
    g = load_golden("loss")
    est, tgt = torch.from_numpy(g["est"]), torch.from_numpy(g["tgt"])
    for name in ("snr", "sisdr", "sdsdr"):
        pw = O.pairwise_neg_sdr(est, tgt, name)
        np.testing.assert_allclose(pw.numpy(), g[f"pw_{name}"], rtol=1e-5, atol=1e-5)
        for thr in (True, False):
            loss, reo = O.pit_loss(est, tgt, name, thr, return_ests=True)
            np.testing.assert_allclose(loss.item(), g[f"pit_{name}_{int(thr)}"], rtol=1e-5, atol=1e-5)
            swapped = (reo[:, 0] == est[:, 1]).all(dim=-1).long().numpy()
            np.testing.assert_array_equal(swapped, g[f"perm0_{name}_{int(thr)}"])
    # threshold_byloss with every item below -30 dB keeps them all
    loss = O.pit_loss(torch.from_numpy(g["est_all_below"]), tgt, "snr", True)
    np.testing.assert_allclose(loss.item(), g["pit_snr_all_below"], rtol=1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/look2hear"), reason="reference tree not present")
def test_oracle_against_live_reference():
    """In the build container, re-run the unmodified reference (fresh seed, different length)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch\n"
        f"sys.path[:0] = ['{root}/oracle/ref_shim', '/root/reference', '{root}']\n"
        "import look2hear.models as RM\n"
        "from oracle import tdanet_oracle as O\n"
        "kw = dict(out_channels=16, in_channels=32, num_blocks=3, upsampling_depth=4, enc_kernel_size=2, num_sources=2)\n"
        "for variant, cls in (('best', 'TDANetBest'), ('fork', 'TDANet')):\n"
        "    torch.manual_seed(5)\n"
        "    m = getattr(RM, cls)(sample_rate=16000, **kw).eval()\n"
        "    x = torch.randn(2, 1, 2501)\n"
        "    with torch.no_grad():\n"
        "        y = m(x); yo = O.forward(m.state_dict(), x, O.OracleConfig(variant=variant, sample_rate=16000, **kw))\n"
        "    assert (y - yo).abs().max() / y.abs().max() < 2e-6, variant\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def golden_drop_masks(g):
    """keep-masks of a *_train fixture -> OracleConfig.drop_masks"""
    masks = []
    for b in range(g["kwargs"]["num_blocks"]):
        m = {}
        for k in ("att", "ao", "f1", "f2", "dp"):
            shape = tuple(int(v) for v in g[f"mshape/{b}/{k}"])
            m[k] = torch.from_numpy(np.unpackbits(g[f"mask/{b}/{k}"])[: int(np.prod(shape))].reshape(shape).copy())
        masks.append(m)
    return masks


@pytest.mark.parametrize("variant", ["best", "fork", "origin"])
def test_train_mode_with_reference_masks(variant):
    """the reference in TRAIN mode (float64; oracle/make_golden_train.py recorded what its nn.Dropout,
    attention-weight dropout and DropPath drew).  With the same keep-masks the oracle reproduces the output and every
    parameter gradient; without them it does not."""
    g = load_golden(f"{variant}_train")
    sd = {k[3:]: torch.from_numpy(g[k]).double().requires_grad_(not k.endswith("pos_enc.pe"))
          for k in g if k.startswith("sd/")}
    x, d, y_ref = (torch.from_numpy(g[k]) for k in ("x", "d", "y"))
    masks = golden_drop_masks(g)
    assert any((m["dp"] == 0).any() for m in masks) and all(0.8 < m["ao"].float().mean() < 0.97 for m in masks)
    cfg = oracle_cfg(variant, g["kwargs"], g["sample_rate"])
    cfg.drop_masks, cfg.dropout, cfg.drop_path = masks, 0.1, 0.1
    y = O.forward(sd, x, cfg)
    assert max_rel(y.detach(), y_ref) < 1e-12
    (y * d).sum().backward()
    n = 0
    for k in g:
        if k.startswith("grad/"):
            assert max_rel(sd[k[5:]].grad, torch.from_numpy(g[k])) < 1e-10, k
            n += 1
    assert n > 60
    with torch.no_grad():
        y_eval = O.forward(sd, x, oracle_cfg(variant, g["kwargs"], g["sample_rate"]))
    assert max_rel(y_eval, y_ref) > 1e-3
