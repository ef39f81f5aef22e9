"""look2hear.metrics mirror for the separation path: SI-SNR / SI-SNRi of an estimate against the clean sources
and against the mixture baseline, on the device through the fused PIT kernel.

Reference: look2hear/metrics/wrapper.py:24-90 (`MetricsTracker`, used by audio_test.py) and the validation /
test steps of `AudioLightningModule` (system/audio_litmodule.py:127-191).  The reference's SDR columns come from
`fast_bss_eval` (512-tap distortion filters, a third-party package that is not part of this path); they are
filled only when that package is importable and left empty otherwise.
"""
import csv

import numpy as np
import torch

from ...engine import pit_loss


def pit_si_snr(est: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    """Best-permutation SI-SNR in dB per item: est, clean [B, n_src, T] (CUDA) -> [B]."""
    _, pw, perm, _ = pit_loss(est, clean, "sisdr", threshold_byloss=False, want_grad=False)
    n = clean.shape[1]
    # mean over targets j of -pw[b, perm[b, j], j]
    picked = torch.gather(pw, 1, perm.long().unsqueeze(1)).squeeze(1)   # [B, n]: pw[b, perm[b, j], j]
    return -picked.mean(dim=1) if n > 0 else picked.sum(dim=1)


def si_snr_improvement(mix: torch.Tensor, clean: torch.Tensor, est: torch.Tensor):
    """(si_snr [B], si_snr_i [B]) with the mixture repeated for every source as the baseline
    (wrapper.py:42-47); mix [B, T], clean / est [B, n_src, T]."""
    s = pit_si_snr(est, clean)
    base = pit_si_snr(mix.unsqueeze(1).expand(-1, clean.shape[1], -1).contiguous(), clean)
    return s, s - base


class MetricsTracker:
    """Per-utterance metric rows + running averages, like the reference class (same CSV columns)."""

    def __init__(self, save_file: str = ""):
        self.all_sdrs, self.all_sdrs_i, self.all_sisnrs, self.all_sisnrs_i = [], [], [], []
        self.results_csv = open(save_file, "w") if save_file else None
        if self.results_csv:
            self.writer = csv.DictWriter(self.results_csv, fieldnames=["snt_id", "sdr", "sdr_i", "si-snr", "si-snr_i"])
            self.writer.writeheader()
        try:
            import fast_bss_eval  # noqa: F401
            self._bss = fast_bss_eval
        except ImportError:
            self._bss = None

    def __call__(self, mix, clean, estimate, key):
        """mix [T], clean / estimate [n_src, T] (one utterance, CUDA tensors)."""
        s, si = si_snr_improvement(mix.unsqueeze(0), clean.unsqueeze(0), estimate.unsqueeze(0))
        row = {"snt_id": key, "sdr": "", "sdr_i": "", "si-snr": s.item(), "si-snr_i": si.item()}
        if self._bss is not None:
            mixr = torch.stack([mix] * clean.shape[0], dim=0)
            sdr = -self._bss.sdr_pit_loss(clean, estimate).mean()
            sdr_i = sdr + self._bss.sdr_pit_loss(mixr, clean).mean()
            row["sdr"], row["sdr_i"] = sdr.item(), sdr_i.item()
            self.all_sdrs.append(row["sdr"])
            self.all_sdrs_i.append(row["sdr_i"])
        if self.results_csv:
            self.writer.writerow(row)
        self.all_sisnrs.append(row["si-snr"])
        self.all_sisnrs_i.append(row["si-snr_i"])
        return row

    def update(self):
        return {"sdr_i": float(np.mean(self.all_sdrs_i)) if self.all_sdrs_i else float("nan"),
                "si-snr_i": float(np.mean(self.all_sisnrs_i))}

    def final(self):
        def agg(fn):
            return {"sdr": fn(self.all_sdrs) if self.all_sdrs else "", "sdr_i": fn(self.all_sdrs_i) if self.all_sdrs_i else "",
                    "si-snr": fn(self.all_sisnrs), "si-snr_i": fn(self.all_sisnrs_i)}
        rows = [dict(snt_id="avg", **agg(lambda v: float(np.mean(v)))), dict(snt_id="std", **agg(lambda v: float(np.std(v))))]
        if self.results_csv:
            for r in rows:
                self.writer.writerow(r)
            self.results_csv.close()
        return rows
