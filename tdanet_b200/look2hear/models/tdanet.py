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


class _SeparateFn(torch.autograd.Function):
    """model(mix) with gradients: tdanet_forward_train / tdanet_backward behind torch.autograd, so that the
    reference's `loss.backward()` keeps working.  (The fused step in look2hear.system.TrainingStep skips
    autograd altogether.)"""

    @staticmethod
    def forward(ctx, model, wav, names, *params):
        ctx.model, ctx.names = model, names
        ctx.save_for_backward(wav, *params)
        model._sync_dropout()
        est = model._engine.forward_train(model._weights(), wav, model.attn_group)
        # the activations live in the engine's per-device training workspace: remember which call wrote them, so
        # that backward() raises if another grad-mode forward (or a train()/eval() switch) came in between
        ctx.generation = model._engine.train_generation(wav.device)
        ctx.attn_group = model.attn_group
        return est

    @staticmethod
    def backward(ctx, d_est):
        model, names = ctx.model, ctx.names
        wav, *params = ctx.saved_tensors
        # gradients land in one zeroed flat buffer; parameters the forward never reads get None, like autograd
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=wav.device)
        views, off = {}, 0
        for n, p, s in zip(names, params, sizes):
            views[n] = flat[off:off + p.numel()].view(p.shape)
            off += s
        gw = model._engine.pack(views, optional=True)
        model._engine.backward(model._weights(), gw, wav, d_est.contiguous().float(), ctx.attn_group,
                               generation=ctx.generation)
        dead = model._unused_parameter_names()
        grads = tuple(None if (n in dead or not p.requires_grad) else views[n] for n, p in zip(names, params))
        return (None, None, None) + grads


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
        # train-mode regularisation, hard-coded in the reference (GA(..., 0.1): DropPath; MultiHeadAttention(.., 0.1),
        # FFN(drop=0.1): nn.Dropout; TDANet_best.py:256-259,335-337).  Applied only while `self.training`.
        self.dropout = 0.1
        self.drop_path = 0.1

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

    def _sync_dropout(self) -> None:
        """Hands the train-mode dropout / DropPath probabilities to the engine (0 / 0 in eval mode, like nn.Dropout)."""
        on = self.training
        self._engine.set_dropout(self.dropout if on else 0.0, self.drop_path if on else 0.0)

    def manual_seed(self, seed: int, device=None) -> None:
        """Seeds the Philox stream of the dropout masks on `device` (default: the parameters' device)."""
        device = device if device is not None else next(self.parameters()).device
        self._engine.rng_state(device, seed)

    def _weights(self):
        # (re)pack when any storage moved (e.g. after .cuda() / load_state_dict with assign)
        sd = {k: v for k, v in self.state_dict(keep_vars=True).items()}
        sig = tuple(t.data_ptr() for t in sd.values())
        if getattr(self, "_packed_sig", None) != sig:
            if getattr(self, "_packed", None) is not None:
                self._engine.drop_graphs(self._packed)     # captured forwards of the old storage
            self._packed = self._engine.pack({k: v.detach() for k, v in sd.items()})
            self._packed_sig = sig
        return self._packed

    def graph_slot(self, batch: int, n_samples: int, slot: int = 0):
        """(graph, static_in [B, T], static_out [B, n_src, T]) of this model's captured inference forward (see
        SeparationEngine.graph_slot); used by look2hear.system.separate_pipelined."""
        dev = next(self.parameters()).device
        return self._engine.graph_slot(self._weights(), batch, n_samples, dev, self.attn_group, slot)

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
        wav = input_wav.float().contiguous()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # gradients w.r.t. the parameters through the hand-written backward pass; in train mode with the
            # reference's dropout / DropPath.  There is no gradient w.r.t. the waveform.
            named = [(n, p) for n, p in self.named_parameters()]
            est = _SeparateFn.apply(self, wav, tuple(n for n, _ in named), *[p for _, p in named])
            return est.squeeze(0) if was_one_d else est
        w = self._weights()
        if self.use_cuda_graph:
            est = self._engine.forward_graphed(w, wav, self.attn_group).clone()
        else:
            est = self._engine.forward(w, wav, self.attn_group)
        return est.squeeze(0) if was_one_d else est

    def _unused_parameter_names(self):
        """Parameters no forward reads (their reference gradient is None): loc_glo_fus of the last scale unless
        it is the first top-down step's partner (depth 2), concat_block when there is a single block."""
        d = self.upsampling_depth
        dead = set()
        if self._variant == "best" and (d - 3 + d) % d != d - 1:
            dead |= {n for n, _ in self.named_parameters() if f"loc_glo_fus.{d - 1}." in n}
        if self.num_blocks == 1:
            dead |= {n for n, _ in self.named_parameters() if "concat_block" in n}
        return dead

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


class TDANetOrigin(_TDANetCommon):
    """The original TDANet (TDANet_origin.py:393): like TDANetBest without loc_glo_fus - average-pool gather,
    additive injection of the global feature, GroupNorm(1, C) norms."""
    _variant = "origin"

    def __init__(self, out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=4,
                 enc_kernel_size=21, num_sources=2, sample_rate=16000, feat_len=3010):
        super().__init__(sample_rate=sample_rate)
        self._build(out_channels, in_channels, num_blocks, upsampling_depth, enc_kernel_size, num_sources,
                    sample_rate)


class TDANetYang(TDANetOrigin):
    """TDANet_yang.py:441: the same computation as TDANetOrigin (its time-axis attention module is defined but not
    used by its GlobalAttention), kept as its own class name for `models.get` / configs/tdanet.yml."""
    _variant = "yang"


class TDANetMultRes(_TDANetCommon):
    """Multi-resolution encoder, no bottleneck, time-axis attention (TDANet_mult_tes.py:455)."""
    _variant = "multres"

    def __init__(self, out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=4,
                 enc_kernel_size=21, num_sources=2, sample_rate=16000, feat_len=3010, kernels=3):
        super().__init__(sample_rate=sample_rate)
        self._build(out_channels, in_channels, num_blocks, upsampling_depth, enc_kernel_size, num_sources,
                    sample_rate, kernels=kernels)
