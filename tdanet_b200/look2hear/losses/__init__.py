"""look2hear.losses mirror (reference: look2hear/losses/__init__.py, matrix.py, pit_wrapper.py)."""
from .pit import (PITLossWrapper, PairwiseNegSDR, pairwise_neg_sdsdr, pairwise_neg_sisdr,
                  pairwise_neg_snr)

__all__ = ["PITLossWrapper", "PairwiseNegSDR", "pairwise_neg_sisdr", "pairwise_neg_sdsdr", "pairwise_neg_snr"]
