"""TDANetBest / TDANet / TDANetMultRes behind the reference's class surface, computed by the CUDA engine.

Reference classes: look2hear/models/TDANet_best.py:402-525, TDANet.py:788-913,
TDANet_mult_tes.py:455-579.  Constructor kwargs, forward(wav) rank handling ([T], [B,T], [B,1,T] ->
[n_src,T] / [B,n_src,T]) and state_dict keys are the same; the arithmetic runs in
tdanet_b200/csrc through the C ABI (include/tdanet_b200.h).  There is no CPU path.
"""
import math

import torch
import torch.nn as nn

from ... import _lib
from ...engine import SeparationEngine
from . import _params as P
from .base_model import BaseModel


class _TDANetCommon(BaseModel):
    _variant = None

    def _build(self, out_channels, in_channels, num_blocks, upsampling_depth, enc_kernel_size, num_sources,
               sample_rate, kernels=1):
        v = self._variant
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_blocks = num_blocks
        self.upsampling_depth = upsampling_depth
        self.enc_kernel_size = enc_kernel_size * sample_rate // 1000
        self.enc_num_basis = self.enc_kernel_size // 2 + 1
        self.num_sources = num_sources
        K = self.enc_kernel_size
        self.lcm = abs(K // 4 * 4 ** upsampling_depth) // math.gcd(K // 4, 4 ** upsampling_depth)
        if v == "multres":
            self.kernels = kernels
            self.encoder = P.ConvEncoderParams(K, out_channels, kernels)
            for conv in self.encoder.conv_list:
                nn.init.xavier_uniform_(conv.weight)
            n_basis = out_channels
            self.ln = P.make_gln(v, out_channels)
        else:
            self.encoder = nn.Conv1d(1, self.enc_num_basis, K, stride=K // 4, padding=K // 2, bias=False)
            nn.init.xavier_uniform_(self.encoder.weight)
            n_basis = self.enc_num_basis
            self.ln = P.make_gln(v, n_basis)
            self.bottleneck = nn.Conv1d(n_basis, out_channels, 1)
        self.sm = P.RecurrentParams(v, out_channels, in_channels, upsampling_depth, 8)
        mask_conv = nn.Conv1d(out_channels, num_sources * n_basis, 1)
        self.mask_net = nn.Sequential(nn.PReLU(), mask_conv)
        self.decoder = nn.ConvTranspose1d(n_basis * num_sources, num_sources, K, stride=K // 4, padding=K // 2,
                                          groups=1, bias=False)
        nn.init.xavier_uniform_(self.decoder.weight)
        self._engine = SeparationEngine(v, out_channels, in_channels, num_blocks, upsampling_depth, K, n_basis,
                                        num_sources, enc_convs=kernels if v == "multres" else 1, n_head=8)
        # knobs of this implementation (not in the reference signature)
        self.attn_group = 0       # 0: the call's whole batch attends together, like the reference
        self.use_cuda_graph = False

    # ------------------------------------------------------------------ engine access
    @property
    def engine(self) -> SeparationEngine:
        return self._engine

    @property
    def gemm_mode(self) -> str:
        return self._engine.gemm_mode

    @gemm_mode.setter
    def gemm_mode(self, mode: str) -> None:
        self._engine.gemm_mode = mode

    @property
    def act_dtype(self) -> str:
        """Storage of the large activations: "fp32", or "bf16" (fp32 arithmetic; the bf16-mode tolerance)."""
        return self._engine.act_dtype

    @act_dtype.setter
    def act_dtype(self, dtype: str) -> None:
        self._engine.act_dtype = dtype

    def _weights(self):
        # (re)pack when any storage moved (e.g. after .cuda() / load_state_dict with assign)
        sd = {k: v for k, v in self.state_dict(keep_vars=True).items()}
        sig = tuple(t.data_ptr() for t in sd.values())
        if getattr(self, "_packed_sig", None) != sig:
            self._packed = self._engine.pack({k: v.detach() for k, v in sd.items()})
            self._packed_sig = sig
        return self._packed

    def pad_input(self, input, window, stride):
        """Kept for API parity (TDANet_best.py:465-479); the CUDA encoder folds this padding into its
        indexing and never materialises it."""
        batch_size, nsample = input.shape
        rest = window - (stride + nsample % window) % window
        if rest > 0:
            input = torch.cat([input, input.new_zeros(batch_size, rest)], 1)
        pad_aux = input.new_zeros(batch_size, window - stride)
        return torch.cat([pad_aux, input, pad_aux], 1), rest

    def forward(self, input_wav):
        was_one_d = input_wav.ndim == 1
        if was_one_d:
            input_wav = input_wav.unsqueeze(0)
        if input_wav.ndim == 3:
            input_wav = input_wav.squeeze(1)
        if not input_wav.is_cuda:
            raise _lib.TdanetError(
                f"{type(self).__name__} (tdanet_b200) runs on CUDA tensors only; there is no CPU fallback")
        if torch.is_grad_enabled() and (input_wav.requires_grad or any(p.requires_grad for p in self.parameters())) \
                and self.training:
            raise NotImplementedError(
                "training-mode forward (autograd through the CUDA path) is not part of this build yet; "
                "call .eval() / torch.no_grad() for separation")
        wav = input_wav.float().contiguous()
        w = self._weights()
        if self.use_cuda_graph:
            est = self._engine.forward_graphed(w, wav, self.attn_group).clone()
        else:
            est = self._engine.forward(w, wav, self.attn_group)
        return est.squeeze(0) if was_one_d else est

    def get_model_args(self):
        return {"n_src": 2}


class TDANetBest(_TDANetCommon):
    """Upstream TDANet ("4 ms encoder, LRS2 architecture"): GlobLN with gamma/beta, average-pool
    gather, gated loc_glo_fus injection (TDANet_best.py:402)."""
    _variant = "best"

    def __init__(self, out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=4,
                 enc_kernel_size=21, num_sources=2, sample_rate=16000):
        super().__init__(sample_rate=sample_rate)
        self._build(out_channels, in_channels, num_blocks, upsampling_depth, enc_kernel_size, num_sources,
                    sample_rate)


class TDANet(_TDANetCommon):
    """Fork variant: learned conv_pool gather, GroupNorm(1, C) norms, additive injection (TDANet.py:788)."""
    _variant = "fork"

    def __init__(self, out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=4,
                 enc_kernel_size=21, num_sources=2, sample_rate=16000, feat_len=None):
        super().__init__(sample_rate=sample_rate)
        self._build(out_channels, in_channels, num_blocks, upsampling_depth, enc_kernel_size, num_sources,
                    sample_rate)


class TDANetMultRes(_TDANetCommon):
    """Multi-resolution encoder, no bottleneck, time-axis attention (TDANet_mult_tes.py:455)."""
    _variant = "multres"

    def __init__(self, out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=4,
                 enc_kernel_size=21, num_sources=2, sample_rate=16000, feat_len=3010, kernels=3):
        super().__init__(sample_rate=sample_rate)
        self._build(out_channels, in_channels, num_blocks, upsampling_depth, enc_kernel_size, num_sources,
                    sample_rate, kernels=kernels)
