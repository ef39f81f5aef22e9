"""Generate tests/golden/*_train.npz: a TRAIN-MODE forward + backward of the UNMODIFIED reference with the
keep-masks its stochastic layers drew.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python oracle/make_golden_train.py

The reference's nn.Dropout / nn.MultiheadAttention(dropout) go through ``torch.nn.functional.dropout`` and its
``drop_path`` through ``torch.rand`` (TDANet_best.py:7-18); both are wrapped here only to RECORD what they drew
(the real functions do the work).  The fixture pins the oracle's explicit-mask restatement of SURVEY.md §8 a21
(``OracleConfig.drop_masks``): same masks in, same output and same parameter gradients out.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TDANET_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

import look2hear.models as RM          # noqa: E402  (the reference)
from oracle import tdanet_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CLASSES = {"best": "TDANetBest", "fork": "TDANet", "origin": "TDANetOrigin"}
SMALL = dict(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=4, enc_kernel_size=4, num_sources=2)


class Recorder:
    """Records, in call order, the keep-mask of every F.dropout call and the uniform draws of every torch.rand call."""

    def __enter__(self):
        self.drop, self.rand = [], []
        self._dropout, self._rand = F.dropout, torch.rand

        def dropout(x, p=0.5, training=True, inplace=False):
            out = self._dropout(x, p, training, False)
            if training and p > 0:
                # where the input is exactly 0 (after the ReLU) the draw is unobservable and irrelevant: call it kept
                self.drop.append(((out != 0) | (x == 0)).detach().clone())
            return out

        def rand(*a, **k):
            r = self._rand(*a, **k)
            self.rand.append(r.detach().clone())
            return r

        F.dropout, torch.rand = dropout, rand
        return self

    def __exit__(self, *exc):
        F.dropout, torch.rand = self._dropout, self._rand


def main():
    report = []
    for variant, cls in CLASSES.items():
        torch.manual_seed(7)
        m = getattr(RM, cls)(sample_rate=8000, **SMALL)
        with torch.no_grad():
            g = torch.Generator().manual_seed(11)
            for k, p in m.named_parameters():
                if p.ndim == 1:
                    p.add_(0.2 * torch.randn(p.shape, generator=g))
        # float64 end to end: in fp32 the autograd chains of the reference and of the oracle differ by their own
        # rounding (up to 1.6e-2 on single tensors of the GroupNorm variant), which would hide a semantic difference
        m.double().train()
        B, T = 5, 1203
        g = torch.Generator().manual_seed(1234)
        x = (torch.randn(B, 1, T, generator=g) * 0.1).double()
        d = torch.randn(B, 2, T, generator=g).double()
        torch.manual_seed(3)
        with Recorder() as rec:
            y = m(x)
        (y * d).sum().backward()
        nb = SMALL["num_blocks"]
        assert len(rec.drop) == 4 * nb and len(rec.rand) == 2 * nb, (len(rec.drop), len(rec.rand))
        masks = []
        for b in range(nb):
            att, ao, f1, f2 = rec.drop[4 * b: 4 * b + 4]
            dp = torch.stack([torch.floor(0.9 + r).flatten() for r in rec.rand[2 * b: 2 * b + 2]])
            masks.append({"att": att.to(torch.uint8), "ao": ao.to(torch.uint8),
                          "f1": f1.transpose(1, 2).contiguous().to(torch.uint8),
                          "f2": f2.transpose(1, 2).contiguous().to(torch.uint8), "dp": dp.to(torch.uint8)})
        assert any((mm["dp"] == 0).any() for mm in masks), "no dropped path in this draw: pick another seed"
        t_bot = masks[0]["ao"].shape[1]
        sd = {k: (v[:, :t_bot] if k.endswith("pos_enc.pe") else v).detach().clone() for k, v in m.state_dict().items()}
        cfg = O.OracleConfig(variant=variant, sample_rate=8000, drop_masks=masks, dropout=0.1, drop_path=0.1, **SMALL)
        sdg = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
        yo = O.forward(sdg, x, cfg)
        (yo * d).sum().backward()
        e_y = (y - yo).abs().max().item() / y.abs().max().item()
        e_g = 0.0
        grads = {}
        for k, p in m.named_parameters():
            if p.grad is None:
                assert sdg[k].grad is None, k
                continue
            grads[k] = p.grad.detach().clone()
            e_g = max(e_g, (p.grad - sdg[k].grad).abs().max().item() / max(p.grad.abs().max().item(), 1e-12))
        report.append(f"{variant}_train (float64): oracle(masks) vs reference train mode: output max-rel {e_y:.3e}, "
                      f"worst parameter-gradient max-rel {e_g:.3e} (B={B}, T={T}, dropped paths "
                      f"{sum(int((mm['dp'] == 0).sum()) for mm in masks)})")
        arrs = {"x": x.numpy(), "d": d.numpy(), "y": y.detach().numpy(), "kwargs": np.array(repr(SMALL)),
                "sample_rate": np.array(8000)}
        arrs.update({"sd/" + k: v.numpy() for k, v in sd.items()})
        arrs.update({"grad/" + k: v.numpy() for k, v in grads.items()})
        for b, mm in enumerate(masks):
            for k, v in mm.items():
                arrs[f"mask/{b}/{k}"] = np.packbits(v.numpy().reshape(-1))
                arrs[f"mshape/{b}/{k}"] = np.array(v.shape)
        np.savez_compressed(os.path.join(OUT, f"{variant}_train.npz"), **arrs)
    with open(os.path.join(OUT, "REPORT.txt"), "a") as f:
        f.write("generated by oracle/make_golden_train.py from the unmodified reference in train mode\n")
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
