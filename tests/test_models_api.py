"""The look2hear-compatible class surface: constructor kwargs, state_dict keys and seeded init
identical to the reference, model registry, from_pretrain round trip, loud failure on CPU input."""
import numpy as np
import pytest
import torch

import tdanet_b200.look2hear as look2hear
from conftest import CLASSES, load_golden
from tdanet_b200 import _lib

M = look2hear.models


@pytest.mark.parametrize("variant", list(CLASSES))
def test_state_dict_and_init_match_reference(variant):
    g = load_golden(f"{variant}_full")
    torch.manual_seed(int(g["init_seed"]))
    m = M.get(CLASSES[variant])(sample_rate=g["sample_rate"], **g["kwargs"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    s1 = np.array([v.double().sum().item() for v in sd.values()])
    s2 = np.array([(v.double() ** 2).sum().item() for v in sd.values()])
    np.testing.assert_allclose(s1, g["sum"], rtol=0, atol=0)
    np.testing.assert_allclose(s2, g["sumsq"], rtol=0, atol=0)
    assert m.get_model_args() == {"n_src": 2} and m.sample_rate() == g["sample_rate"]


def test_registry():
    assert M.get("tdanetbest") is M.TDANetBest and M.get("TDANET") is M.TDANet
    with pytest.raises(ValueError):
        M.get("nope")
    with pytest.raises(ValueError):
        M.register_model(M.TDANet)


def test_from_pretrain_round_trip(tmp_path):
    kw = dict(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=4, enc_kernel_size=2, num_sources=2)
    torch.manual_seed(3)
    m = M.TDANetBest(sample_rate=16000, **kw)
    # Lightning-style checkpoint: keys prefixed with audio_model.
    ckpt = {"state_dict": {"audio_model." + k: v for k, v in m.state_dict().items()}}
    path = tmp_path / "ckpt.pth"
    torch.save(ckpt, path)
    m2 = M.BaseModel.from_pretrain("TDANetBest", str(path), sample_rate=16000, **kw)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    conf = m.serialize()
    assert conf["model_name"] == "TDANetBest" and conf["model_args"] == {"n_src": 2}


def test_cpu_input_fails_loudly():
    m = M.TDANetBest(out_channels=16, in_channels=32, num_blocks=1, upsampling_depth=4, enc_kernel_size=2).eval()
    with pytest.raises(_lib.TdanetError, match="CUDA"):
        m(torch.zeros(1, 1, 800))
    with pytest.raises(_lib.TdanetError, match="CUDA"):
        look2hear.losses.PITLossWrapper(look2hear.losses.pairwise_neg_snr)(torch.zeros(1, 2, 100), torch.zeros(1, 2, 100))


def test_loss_surface():
    L = look2hear.losses
    assert L.pairwise_neg_sisdr.sdr_type == "sisdr"
    with pytest.raises(ValueError):
        L.PITLossWrapper(L.pairwise_neg_snr, pit_from="bogus")
    with pytest.raises(NotImplementedError):
        L.PITLossWrapper(L.pairwise_neg_snr, pit_from="perm_avg")
