"""The C-ABI library: it loads without a GPU, exports every symbol include/tdanet_b200.h declares,
its structs match the ctypes mirrors, and host-side planning/validation behaves.  CPU only."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_golden, oracle_cfg
from tdanet_b200 import _lib
from tdanet_b200.engine import SeparationEngine


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tdanet_b200.h")).read()
    declared = set(re.findall(r"TDANET_API[^;(]*?\b(tdanet_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert declared == set(_lib.EXPORTS)
    assert lib.tdanet_abi_version() == 3


def test_struct_sizes_match():
    lib = _lib.load()
    cb, wb = C.c_size_t(), C.c_size_t()
    assert lib.tdanet_abi_sizes(C.byref(cb), C.byref(wb)) == 0
    assert cb.value == C.sizeof(_lib.Config) == 64
    assert wb.value == C.sizeof(_lib.Weights)


@pytest.mark.parametrize("variant", ["best", "fork", "multres"])
def test_latent_lengths_match_oracle(variant):
    import torch
    from oracle import tdanet_oracle as O
    from conftest import golden_state_dict
    g = load_golden(f"{variant}_small")
    kw, sr = g["kwargs"], g["sample_rate"]
    K = kw["enc_kernel_size"] * sr // 1000
    nb = kw["out_channels"] if variant == "multres" else K // 2 + 1
    eng = SeparationEngine(variant, kw["out_channels"], kw["in_channels"], kw["num_blocks"], kw["upsampling_depth"],
                           K, nb, 2, enc_convs=kw.get("kernels", 1))
    T = g["x"].shape[-1]
    lens, tp, rest = eng.latent_lengths(T)
    cfg = oracle_cfg(variant, kw, sr, taps={})
    with torch.no_grad():
        O.forward(golden_state_dict(g), torch.from_numpy(g["x"])[:1], cfg)
    assert lens[0] == cfg.taps["enc"].shape[-1]
    assert lens[-1] == cfg.taps["spp.4"].shape[-1]
    xp, r = O.pad_input(torch.zeros(1, T), K, K // 4)
    assert (tp, rest) == (xp.shape[1], r)
    assert eng.workspace_bytes(2, T) > eng.workspace_bytes(1, T) > 0


def test_bad_configurations_are_rejected():
    eng = SeparationEngine("best", 128, 512, 16, 5, 64, 33, 2)
    eng.cfg.n_basis = 34
    with pytest.raises(_lib.TdanetError, match="n_basis"):
        eng.workspace_bytes(1, 32000)
    eng = SeparationEngine("best", 100, 512, 16, 5, 64, 33, 2)
    with pytest.raises(_lib.TdanetError, match="out_channels"):
        eng.workspace_bytes(1, 32000)
    eng = SeparationEngine("best", 128, 512, 16, 5, 64, 33, 2)
    with pytest.raises(_lib.TdanetError):
        eng.workspace_bytes(0, 32000)
    lib = _lib.load()
    assert lib.tdanet_workspace_tensor(C.byref(eng.cfg), 1, 32000, b"no_such_tensor", None, None) != 0
    assert b"no_such_tensor" in lib.tdanet_last_error()


def test_headline_shapes():
    eng = SeparationEngine("best", 128, 512, 16, 5, 64, 33, 2)
    assert eng.latent_lengths(32000) == ([2010, 1005, 503, 252, 126], 32144, 48)
    eng2 = SeparationEngine("best", 128, 512, 16, 5, 32, 17, 2)
    assert eng2.latent_lengths(32000)[0] == [4010, 2005, 1003, 502, 251]


def test_deterministic_mode_switch_and_workspace():
    """tdanet_set_deterministic is process-wide, off by default, and enlarges the inference workspace by four bytes
    per byte of GlobLN statistics (the exact accumulators) - the training workspace is not affected."""
    eng = SeparationEngine("best", 128, 512, 16, 5, 64, 33, 2)
    assert not _lib.deterministic()
    plain, train = eng.workspace_bytes(64, 32000), eng.train_workspace_bytes(2, 8000)
    _lib.set_deterministic(True)
    try:
        assert _lib.deterministic()
        det = eng.workspace_bytes(64, 32000)
        # per block: 3 + 4*depth + depth(la_g counts double) item sums [B,2] double, (depth + 1) channel tables [B,2,C]
        stats = 64 * 2 * 512 * 4 * 6
        assert plain + 4 * stats < det < plain + 4 * stats + (1 << 20)
        assert eng.train_workspace_bytes(2, 8000) == train
    finally:
        _lib.set_deterministic(False)
    assert eng.workspace_bytes(64, 32000) == plain
