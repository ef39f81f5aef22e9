"""PairwiseNegSDR / PITLossWrapper on the fused CUDA kernel (csrc/loss.cu).

Reference: look2hear/losses/matrix.py:12-56, 176-178 and pit_wrapper.py:14-131.  Supported: the
configuration every reference YAML uses - PITLossWrapper(pairwise_neg_{snr,sisdr,sdsdr},
pit_from="pw_mtx", perm_reduce=None) with n_src <= 3.  The Hungarian path (n_src > 3), `pw_pt` /
`perm_avg` and custom `perm_reduce` are outside the hot path and raise.
"""
import torch
from torch import nn

from ...engine import pit_loss


class PairwiseNegSDR(nn.Module):
    def __init__(self, sdr_type, zero_mean=True, take_log=True, EPS=1e-8):
        super().__init__()
        assert sdr_type in ["snr", "sisdr", "sdsdr"]
        if not (zero_mean and take_log and EPS == 1e-8):
            raise NotImplementedError("the CUDA kernel implements zero_mean=True, take_log=True, EPS=1e-8")
        self.sdr_type = sdr_type
        self.zero_mean, self.take_log, self.EPS = zero_mean, take_log, EPS

    def forward(self, ests, targets):
        if targets.size() != ests.size() or targets.ndim != 3:
            raise TypeError(
                f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {ests.size()} instead")
        _, pw, _, _ = pit_loss(ests.detach(), targets.detach(), self.sdr_type, False, want_grad=False)
        return pw


class _PITFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ests, targets, sdr_type, threshold):
        loss, _, perm, grad = pit_loss(ests, targets, sdr_type, threshold, want_grad=ests.requires_grad)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(perm)
        return loss.reshape(()), perm

    @staticmethod
    def backward(ctx, g_loss, _g_perm):
        (grad,) = ctx.saved_tensors
        return (None if grad is None else grad * g_loss), None, None, None


class PITLossWrapper(nn.Module):
    def __init__(self, loss_func, pit_from="pw_mtx", perm_reduce=None, threshold_byloss=True):
        super().__init__()
        if pit_from not in ["pw_mtx", "pw_pt", "perm_avg"]:
            raise ValueError(
                "Unsupported loss function type {} for now. Expected"
                "one of [`pw_mtx`, `pw_pt`, `perm_avg`]".format(pit_from))
        if pit_from != "pw_mtx" or perm_reduce is not None or not isinstance(loss_func, PairwiseNegSDR):
            raise NotImplementedError(
                "only PITLossWrapper(PairwiseNegSDR(...), pit_from='pw_mtx', perm_reduce=None) is on the hot path")
        self.loss_func = loss_func
        self.pit_from = pit_from
        self.perm_reduce = perm_reduce
        self.threshold_byloss = threshold_byloss

    def forward(self, ests, targets, return_ests=False, reduce_kwargs=None, **kwargs):
        n_src = targets.shape[1]
        if n_src > 3:
            raise NotImplementedError("n_src > 3 uses the Hungarian solver on the host; outside the hot path")
        loss, perm = _PITFunction.apply(ests, targets, self.loss_func.sdr_type, self.threshold_byloss)
        if not return_ests:
            return loss
        return loss, self.reordered_sources(ests, perm.long())

    @staticmethod
    def reordered_sources(sources, batch_indices):
        idx = batch_indices.unsqueeze(-1).expand(-1, -1, sources.shape[-1])
        return torch.gather(sources, 1, idx)


pairwise_neg_sisdr = PairwiseNegSDR("sisdr")
pairwise_neg_sdsdr = PairwiseNegSDR("sdsdr")
pairwise_neg_snr = PairwiseNegSDR("snr")
