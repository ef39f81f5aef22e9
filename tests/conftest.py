import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLASSES = {"best": "TDANetBest", "fork": "TDANet", "multres": "TDANetMultRes", "origin": "TDANetOrigin",
           "yang": "TDANetYang"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k: g[k] for k in g.files}
    if "kwargs" in out:
        out["kwargs"] = ast.literal_eval(str(out["kwargs"]))
        out["sample_rate"] = int(out["sample_rate"])
    return out


def golden_state_dict(g):
    return {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}


def oracle_cfg(variant, kwargs, sample_rate, **extra):
    from oracle import tdanet_oracle as O
    kw = {k: v for k, v in kwargs.items() if k != "feat_len"}
    return O.OracleConfig(variant=variant, sample_rate=sample_rate, **kw, **extra)


def max_rel(a, b):
    """max |a-b| / max |b|  (the acceptance metric of SURVEY.md §8c)"""
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)
