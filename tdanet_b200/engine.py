"""Host-side driver of the C-ABI library: packs a module's parameters into `tdanet_weights_t`,
owns the workspace, enqueues `tdanet_forward` on the current CUDA stream, and (optionally) replays
the whole forward as one CUDA graph.

PyTorch is used for device memory and streams only; every kernel is in csrc/.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import Config, Weights, check


def _ptr(t: Optional[torch.Tensor], allow_host: bool = False) -> Optional[int]:
    if t is None:
        return None
    if allow_host and not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
        return t.data_ptr()   # tests/test_backward_emu.py only: the CPU emulation build of the kernels
    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.TdanetError(
            f"parameter must be a contiguous fp32 CUDA tensor (got {t.dtype}, {t.device}); "
            "move the model with .cuda() - there is no CPU path")
    return t.data_ptr()


class SeparationEngine:
    """One model configuration bound to the CUDA library.

    variant        "best" | "fork" | "multres" (TDANetBest / TDANet / TDANetMultRes)
    gemm_mode      "fp32" (CUDA cores, exact parity) | "tf32" (tcgen05) | "tf32x3" (tcgen05, split weights)
    """

    def __init__(self, variant: str, out_channels: int, in_channels: int, num_blocks: int, depth: int,
                 enc_kernel: int, n_basis: int, num_sources: int, enc_convs: int = 1, n_head: int = 8,
                 gemm_mode: str = "tf32", act_dtype: str = "fp32"):
        self.variant = variant
        self.cfg = Config(
            variant=_lib.VARIANTS[variant], out_channels=out_channels, in_channels=in_channels,
            num_blocks=num_blocks, depth=depth, enc_kernel=enc_kernel, enc_stride=enc_kernel // 4,
            n_basis=n_basis, num_sources=num_sources, enc_convs=enc_convs, n_head=n_head,
            gemm_mode=_lib.GEMM_MODES[gemm_mode], attn_group=0, act_dtype=_lib.ACT_DTYPES[act_dtype])
        self._ws: Dict[torch.device, torch.Tensor] = {}
        self._tws: Dict[torch.device, torch.Tensor] = {}   # training workspace (activations of every block)
        self._graphs: Dict[Tuple, Tuple] = {}
        self._tws_gen: Dict[torch.device, Tuple] = {}      # what the last forward_train left there (see backward)
        self._gen = 0
        self._tws_listeners = []                           # called when a training workspace is reallocated
        self._rng: Dict[torch.device, torch.Tensor] = {}   # device int64[2] = {seed, offset} of the dropout masks
        self._pack_seq = 0                                 # id of the last pack(): the graph cache is keyed on it

    # ------------------------------------------------------------------ configuration
    @property
    def gemm_mode(self) -> str:
        return {v: k for k, v in _lib.GEMM_MODES.items()}[self.cfg.gemm_mode]

    @gemm_mode.setter
    def gemm_mode(self, mode: str) -> None:
        self.cfg.gemm_mode = _lib.GEMM_MODES[mode]
        self._graphs.clear()

    @property
    def act_dtype(self) -> str:
        return {v: k for k, v in _lib.ACT_DTYPES.items()}[self.cfg.act_dtype]

    @act_dtype.setter
    def act_dtype(self, dtype: str) -> None:
        self.cfg.act_dtype = _lib.ACT_DTYPES[dtype]
        self._graphs.clear()

    def set_dropout(self, dropout: float, drop_path: float) -> None:
        """Training-mode probabilities of forward_train / backward (the reference hard-codes 0.1 / 0.1,
        TDANet_best.py:256-259,335-337); 0 / 0 is the deterministic form every parity test uses."""
        self.cfg.dropout, self.cfg.drop_path = float(dropout), float(drop_path)

    def rng_state(self, device, seed: Optional[int] = None) -> torch.Tensor:
        """Device int64[2] = {seed, offset} that tdanet_forward_train_rng reads and advances.  Seeded from
        torch's default generator on first use (so torch.manual_seed makes the masks reproducible)."""
        device = torch.device(device)
        if seed is not None or device not in self._rng:
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self._rng[device] = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        return self._rng[device]

    def latent_lengths(self, n_samples: int):
        lens = (C.c_int32 * _lib.MAX_DEPTH)()
        tp, rest = C.c_int32(), C.c_int32()
        check(_lib.load().tdanet_latent_lengths(C.byref(self.cfg), n_samples, C.byref(lens), C.byref(tp), C.byref(rest)))
        return list(lens[: self.cfg.depth]), tp.value, rest.value

    # ------------------------------------------------------------------ weights
    def pack(self, sd: Dict[str, torch.Tensor], _allow_host: bool = False, optional: bool = False) -> Weights:
        """state_dict (reference key names) -> tdanet_weights_t of device pointers.

        `optional`: missing keys become NULL (used to pack gradient buffers, which have no `pe`)."""
        v, d = self.variant, self.cfg.depth
        gk, bk = ("gamma", "beta") if v == "best" else ("weight", "bias")
        w = Weights()
        keep = []

        all_optional = optional

        def P(key, optional=False):
            t = sd.get(key)
            if t is None:
                if optional or all_optional:
                    return None
                raise KeyError(f"state_dict has no '{key}'")
            keep.append(t)
            return _ptr(t, _allow_host)

        def convnorm(dst, prefix, bias):
            dst.w = P(f"{prefix}.conv.weight")
            dst.b = P(f"{prefix}.conv.bias") if bias else None
            dst.gamma = P(f"{prefix}.norm.{gk}")
            dst.beta = P(f"{prefix}.norm.{bk}")

        def la(dst, prefix):
            convnorm(dst.local_embedding, f"{prefix}.local_embedding", False)
            convnorm(dst.global_embedding, f"{prefix}.global_embedding", False)
            convnorm(dst.global_act, f"{prefix}.global_act", False)

        if v == "multres":
            for k in range(self.cfg.enc_convs):
                w.enc_w[k] = P(f"encoder.conv_list.{k}.weight")
        else:
            w.enc_w[0] = P("encoder.weight")
            w.bottleneck_w, w.bottleneck_b = P("bottleneck.weight"), P("bottleneck.bias")
        w.ln_gamma, w.ln_beta = P(f"ln.{gk}"), P(f"ln.{bk}")
        u = "sm.unet"
        convnorm(w.proj, f"{u}.proj_1x1", True)
        w.proj_prelu = P(f"{u}.proj_1x1.act.weight")
        for k in range(d):
            convnorm(w.spp_dw[k], f"{u}.spp_dw.{k}", True)
            if v == "best":
                la(w.loc_glo_fus[k], f"{u}.loc_glo_fus.{k}")
            if v == "fork":
                q, cp = f"{u}.conv_pool.{k}", w.conv_pool[k]
                cp.dw_w, cp.dw_b = P(f"{q}.dw_conv.weight"), P(f"{q}.dw_conv.bias")
                cp.pw_w, cp.pw_b = P(f"{q}.pw_conv.weight"), P(f"{q}.pw_conv.bias")
                cp.gamma, cp.beta = P(f"{q}.norm.{gk}"), P(f"{q}.norm.{bk}")
        for i in range(d - 1):
            la(w.last_layer[i], f"{u}.last_layer.{i}")
        w.res_w, w.res_b = P(f"{u}.res_conv.weight"), P(f"{u}.res_conv.bias")
        a = f"{u}.globalatt.attn"
        pe = sd.get(f"{a}.pos_enc.pe")
        w.pe, w.pe_rows = P(f"{a}.pos_enc.pe"), (int(pe.shape[1]) if pe is not None else 0)
        w.ln1_w, w.ln1_b = P(f"{a}.attn_in_norm.weight"), P(f"{a}.attn_in_norm.bias")
        w.in_proj_w, w.in_proj_b = P(f"{a}.attn.in_proj_weight"), P(f"{a}.attn.in_proj_bias")
        w.out_proj_w, w.out_proj_b = P(f"{a}.attn.out_proj.weight"), P(f"{a}.attn.out_proj.bias")
        w.ln2_w, w.ln2_b = P(f"{a}.norm.weight"), P(f"{a}.norm.bias")
        m = f"{u}.globalatt.mlp"
        convnorm(w.fc1, f"{m}.fc1", False)
        w.ffn_dw_w, w.ffn_dw_b = P(f"{m}.dwconv.weight"), P(f"{m}.dwconv.bias")
        convnorm(w.fc2, f"{m}.fc2", False)
        w.concat_w, w.concat_b = P("sm.concat_block.0.weight"), P("sm.concat_block.0.bias")
        w.concat_prelu = P("sm.concat_block.1.weight")
        w.mask_prelu, w.mask_w, w.mask_b = P("mask_net.0.weight"), P("mask_net.1.weight"), P("mask_net.1.bias")
        w.dec_w = P("decoder.weight")
        # The struct holds raw pointers: the tensors must outlive it.  They ride on the struct itself, so whoever owns
        # the Weights (the model's cached pack, a TrainingStep, one autograd backward) owns their lifetime - nothing
        # accumulates on the engine.
        w._keep = keep
        # Captured graphs bake these pointers in.  They are cached per pack id, not per struct address: a struct
        # garbage-collected after a re-pack can hand its address to a later one, and a graph of the freed storage
        # must never be replayed for it.
        self._pack_seq += 1
        w._pack_id = self._pack_seq
        return w

    def drop_graphs(self, weights: Optional[Weights] = None) -> None:
        """Forgets the captured forwards of `weights` (all of them when None); the model calls this when it re-packs
        its parameters after their storage moved."""
        if weights is None:
            self._graphs.clear()
            return
        pid = getattr(weights, "_pack_id", None)
        for key in [k for k in self._graphs if k[6] == pid]:
            del self._graphs[key]

    # ------------------------------------------------------------------ workspace
    def workspace_bytes(self, batch: int, n_samples: int) -> int:
        n = C.c_size_t()
        check(_lib.load().tdanet_workspace_bytes(C.byref(self.cfg), batch, n_samples, C.byref(n)))
        return n.value

    def _workspace(self, device, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(device)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws[device] = ws
            self._graphs.clear()
        return ws

    def workspace_tensor(self, name: str, batch: int, n_samples: int, device) -> torch.Tensor:
        """View of an intermediate of the last forward (channels-last [B, L, C]); for tests."""
        off, dims = C.c_size_t(), (C.c_int64 * 3)()
        check(_lib.load().tdanet_workspace_tensor(C.byref(self.cfg), batch, n_samples, name.encode(), C.byref(off), C.byref(dims)))
        ws = self._ws[torch.device(device)]
        n = dims[0] * dims[1] * dims[2]
        return ws[off.value: off.value + 4 * n].view(torch.float32).view(dims[0], dims[1], dims[2])

    # ------------------------------------------------------------------ forward
    def forward(self, weights: Weights, wav: torch.Tensor, attn_group: int = 0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """wav [B, T] fp32 CUDA contiguous -> est [B, num_sources, T]; enqueued on the current stream."""
        if not wav.is_cuda:
            raise _lib.TdanetError("tdanet_b200 runs on CUDA tensors only (no CPU path); got a CPU tensor")
        if wav.dtype != torch.float32 or wav.ndim != 2:
            raise _lib.TdanetError(f"wav must be fp32 [B, T], got {wav.dtype} {tuple(wav.shape)}")
        wav = wav.contiguous()
        B, T = wav.shape
        lib = _lib.load()
        with torch.cuda.device(wav.device):
            nbytes = self.workspace_bytes(B, T)
            ws = self._workspace(wav.device, nbytes)
            if out is None:
                out = torch.empty(B, self.cfg.num_sources, T, dtype=torch.float32, device=wav.device)
            self.cfg.attn_group = attn_group
            stream = torch.cuda.current_stream(wav.device).cuda_stream
            check(lib.tdanet_forward(C.byref(self.cfg), C.byref(weights), wav.data_ptr(), B, T, out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), stream))
        return out

    # ------------------------------------------------------------------ training: forward keeping activations, backward
    def train_workspace_bytes(self, batch: int, n_samples: int) -> int:
        n = C.c_size_t()
        check(_lib.load().tdanet_train_workspace_bytes(C.byref(self.cfg), batch, n_samples, C.byref(n)))
        return n.value

    def _train_workspace(self, device, nbytes: int) -> torch.Tensor:
        ws = self._tws.get(device)
        if ws is None or ws.numel() < nbytes:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.TdanetError("the training workspace would be reallocated during stream capture")
            ws = None
            self._tws.pop(device, None)
            for cb in list(self._tws_listeners):   # captured training graphs hold the old pointer: drop them
                cb(device)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._tws[device] = ws
        return ws

    def on_train_workspace_realloc(self, callback) -> None:
        """`callback(device)` runs before the training workspace of `device` is replaced by a larger one
        (TrainingStep drops its captured CUDA graph there: the graph bakes the workspace pointer in)."""
        self._tws_listeners.append(callback)

    def _train_signature(self, B: int, T: int, attn_group: int) -> Tuple:
        c = self.cfg
        return (B, T, attn_group, c.gemm_mode, c.act_dtype, float(c.dropout), float(c.drop_path))

    def train_workspace_tensor(self, name: str, block: int, batch: int, n_samples: int, device) -> torch.Tensor:
        """View of a tensor kept by forward_train for UConvBlock iteration `block` (tests)."""
        off, dims, es = C.c_size_t(), (C.c_int64 * 3)(), C.c_int32()
        check(_lib.load().tdanet_train_workspace_tensor(C.byref(self.cfg), batch, n_samples, name.encode(), block,
                                                        C.byref(off), C.byref(dims), C.byref(es)))
        ws = self._tws[torch.device(device)]
        n = dims[0] * dims[1] * dims[2]
        dt = {1: torch.uint8, 2: torch.bfloat16, 4: torch.float32, 8: torch.float64}[es.value]
        return ws[off.value: off.value + es.value * n].view(dt).view(dims[0], dims[1], dims[2])

    def _check_wav(self, wav):
        if not wav.is_cuda:
            raise _lib.TdanetError("tdanet_b200 runs on CUDA tensors only (no CPU path); got a CPU tensor")
        if wav.dtype != torch.float32 or wav.ndim != 2:
            raise _lib.TdanetError(f"wav must be fp32 [B, T], got {wav.dtype} {tuple(wav.shape)}")

    def forward_train(self, weights: Weights, wav: torch.Tensor, attn_group: int = 0,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Like forward(), but every block's activations stay in the training workspace for backward()."""
        self._check_wav(wav)
        wav = wav.contiguous()
        B, T = wav.shape
        lib = _lib.load()
        with torch.cuda.device(wav.device):
            ws = self._train_workspace(wav.device, self.train_workspace_bytes(B, T))
            if out is None:
                out = torch.empty(B, self.cfg.num_sources, T, dtype=torch.float32, device=wav.device)
            self.cfg.attn_group = attn_group
            stream = torch.cuda.current_stream(wav.device).cuda_stream
            if self.cfg.dropout > 0 or self.cfg.drop_path > 0:
                rng = self.rng_state(wav.device)
                check(lib.tdanet_forward_train_rng(C.byref(self.cfg), C.byref(weights), wav.data_ptr(), B, T,
                                                   out.data_ptr(), ws.data_ptr(), ws.numel(), rng.data_ptr(), stream))
            else:
                check(lib.tdanet_forward_train(C.byref(self.cfg), C.byref(weights), wav.data_ptr(), B, T, out.data_ptr(),
                                               ws.data_ptr(), ws.numel(), stream))
            # stamp what now lives in the workspace: backward() refuses anything else (a second
            # forward_train before the first backward used to hand the first graph the second call's activations)
            self._gen += 1
            self._tws_gen[wav.device] = (self._gen, self._train_signature(B, T, attn_group))
        return out

    def train_generation(self, device) -> int:
        """Id of the forward_train call whose activations the training workspace of `device` holds (0: none)."""
        return self._tws_gen.get(torch.device(device), (0, None))[0]

    def backward(self, weights: Weights, grad_weights: Weights, wav: torch.Tensor, d_est: torch.Tensor,
                 attn_group: int = 0, generation: Optional[int] = None) -> None:
        """Adds d loss / d theta into the buffers `grad_weights` points at, from d_est = d loss / d est and the
        workspace the matching forward_train() call left behind.  `generation` (train_generation() right after
        that forward) makes the match explicit: a later forward_train on the same device has overwritten the
        activations and the call raises instead of returning gradients of the wrong graph."""
        self._check_wav(wav)
        B, T = wav.shape
        if d_est.shape != (B, self.cfg.num_sources, T) or d_est.dtype != torch.float32 or not d_est.is_cuda:
            raise _lib.TdanetError(f"d_est must be fp32 CUDA [B, n_src, T], got {d_est.dtype} {tuple(d_est.shape)}")
        d_est = d_est.contiguous()
        lib = _lib.load()
        with torch.cuda.device(wav.device):
            ws = self._tws.get(wav.device)
            need = self.train_workspace_bytes(B, T)
            if ws is None or ws.numel() < need:
                raise _lib.TdanetError("backward() without a matching forward_train() on this device")
            gen, sig = self._tws_gen.get(wav.device, (0, None))
            if generation is not None and generation != gen:
                raise _lib.TdanetError(
                    f"backward() of forward_train call #{generation}, but the training workspace holds call #{gen}: "
                    "a later model(x) in grad mode overwrote the activations (one forward per backward per device)")
            if sig != self._train_signature(B, T, attn_group):
                raise _lib.TdanetError(
                    f"backward() with (B, T, attn_group, gemm_mode, act_dtype, dropout, drop_path) = "
                    f"{self._train_signature(B, T, attn_group)} but the workspace was written with {sig}")
            self.cfg.attn_group = attn_group
            stream = torch.cuda.current_stream(wav.device).cuda_stream
            check(lib.tdanet_backward(C.byref(self.cfg), C.byref(weights), C.byref(grad_weights), wav.data_ptr(),
                                      d_est.data_ptr(), B, T, ws.data_ptr(), ws.numel(), stream))

    def graph_slot(self, weights: Weights, B: int, T: int, device, attn_group: int = 0, slot: int = 0,
                   example: Optional[torch.Tensor] = None):
        """(graph, static_in [B, T], static_out [B, n_src, T]) of the captured forward for this shape; one capture per
        (shape, modes, weights, slot).  Slots are independent input / output buffers over the same workspace - a
        pipeline alternates two of them so that it can copy into / out of the buffers of one step while the other
        step's replay runs (look2hear.system.separate_pipelined), with no staging copies."""
        device = torch.device(device)
        pack_id = getattr(weights, "_pack_id", None)
        if pack_id is None:                  # a struct filled by hand: give it an id of its own
            self._pack_seq += 1
            pack_id = weights._pack_id = self._pack_seq
        key = (device, B, T, attn_group, self.cfg.gemm_mode, self.cfg.act_dtype, pack_id, slot, _lib.deterministic())
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.zeros(B, T, dtype=torch.float32, device=device) if example is None else example.clone()
            static_out = torch.empty(B, self.cfg.num_sources, T, dtype=torch.float32, device=device)
            s = torch.cuda.Stream(device)
            s.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(s):   # warm-up outside capture (function attributes, workspace)
                self.forward(weights, static_in, attn_group, out=static_out)
            torch.cuda.current_stream(device).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward(weights, static_in, attn_group, out=static_out)
            entry = (g, static_in, static_out, self._ws[device])
            self._graphs[key] = entry
        return entry[0], entry[1], entry[2]

    def forward_graphed(self, weights: Weights, wav: torch.Tensor, attn_group: int = 0) -> torch.Tensor:
        """Same as forward() but replays one captured CUDA graph per (B, T, group, weights) key.

        The returned tensor is the graph's static output buffer: it is overwritten by the next call.
        """
        B, T = wav.shape
        g, static_in, static_out = self.graph_slot(weights, B, T, wav.device, attn_group, 0, example=wav)
        static_in.copy_(wav)
        g.replay()
        return static_out


# ---------------------------------------------------------------------- standalone ops
def gemm(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, mode: str = "tf32",
         with_stats: bool = False):
    """D[b, r, :] = A[b, r, :] @ W.T (+ bias) through tdanet_gemm; A [B, L, K], W [N, K]."""
    lib = _lib.load()
    B, L, K = A.shape
    N = W.shape[0]
    D = torch.empty(B, L, N, dtype=torch.float32, device=A.device)
    stats = torch.zeros(B, 2, dtype=torch.float64, device=A.device) if with_stats else None
    nws = lib.tdanet_gemm_workspace_bytes(N, K)
    ws = torch.empty(nws, dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        check(lib.tdanet_gemm(_lib.GEMM_MODES[mode], A.contiguous().data_ptr(), W.contiguous().data_ptr(),
                              None if bias is None else bias.data_ptr(), D.data_ptr(), B, L, N, K,
                              None if stats is None else stats.data_ptr(), ws.data_ptr(), nws,
                              torch.cuda.current_stream(A.device).cuda_stream))
    return (D, stats) if with_stats else D


def pit_loss(est: torch.Tensor, tgt: torch.Tensor, sdr_type: str = "snr", threshold_byloss: bool = True,
             want_grad: bool = True):
    """Fused PIT loss forward (+ d loss / d est).  Returns (loss[1], pw[B,n,n], perm[B,n] int32, grad | None)."""
    if not (est.is_cuda and tgt.is_cuda):
        raise _lib.TdanetError("pit_loss runs on CUDA tensors only (no CPU path)")
    if est.shape != tgt.shape or est.ndim != 3:
        raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {tuple(tgt.shape)} and {tuple(est.shape)} instead")
    lib = _lib.load()
    est = est.contiguous().float()
    tgt = tgt.contiguous().float()
    B, n_src, T = est.shape
    dev = est.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    pw = torch.empty(B, n_src, n_src, dtype=torch.float32, device=dev)
    perm = torch.empty(B, n_src, dtype=torch.int32, device=dev)
    grad = torch.empty_like(est) if want_grad else None
    ns = lib.tdanet_pit_loss_scratch_bytes(B, n_src)
    scratch = torch.empty(ns, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.tdanet_pit_loss(est.data_ptr(), tgt.data_ptr(), B, n_src, T, _lib.SDR_TYPES[sdr_type],
                                  int(bool(threshold_byloss)), loss.data_ptr(), pw.data_ptr(), perm.data_ptr(),
                                  None if grad is None else grad.data_ptr(), scratch.data_ptr(), ns,
                                  torch.cuda.current_stream(dev).cuda_stream))
    return loss, pw, perm, grad


def grad_sqnorm(flat_grads: torch.Tensor, out: torch.Tensor) -> None:
    """out[0] = sum of squares of a flat fp32 gradient buffer (out: float64[2], CUDA)."""
    lib = _lib.load()
    with torch.cuda.device(flat_grads.device):
        check(lib.tdanet_grad_sqnorm(flat_grads.data_ptr(), flat_grads.numel(), out.data_ptr(),
                                     torch.cuda.current_stream(flat_grads.device).cuda_stream))


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
              step: torch.Tensor, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: float = 0.0,
              grad_scale: float = 1.0, sqnorm: Optional[torch.Tensor] = None) -> None:
    """clip_grad_norm_(max_grad_norm) + torch.optim.Adam step on flat fp32 CUDA buffers (csrc/optim.cu)."""
    lib = _lib.load()
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()):
            raise _lib.TdanetError("adam_step needs flat contiguous fp32 CUDA buffers of equal length")
    with torch.cuda.device(params.device):
        check(lib.tdanet_adam_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                   params.numel(), lr, betas[0], betas[1], eps, max_grad_norm, grad_scale,
                                   None if sqnorm is None else sqnorm.data_ptr(), step.data_ptr(),
                                   torch.cuda.current_stream(params.device).cuda_stream))


def wgrad(G: torch.Tensor, A: torch.Tensor, mode: str = "tf32", with_bias: bool = True):
    """dW = G^T A ([N, K]) and db = column sums of G through tdanet_wgrad; G [R, N], A [R, K] fp32 CUDA."""
    lib = _lib.load()
    R, N = G.shape
    K = A.shape[1]
    dW = torch.zeros(N, K, dtype=torch.float32, device=G.device)
    db = torch.zeros(N, dtype=torch.float32, device=G.device) if with_bias else None
    with torch.cuda.device(G.device):
        check(lib.tdanet_wgrad(_lib.GEMM_MODES[mode], G.contiguous().data_ptr(), A.contiguous().data_ptr(), dW.data_ptr(),
                               None if db is None else db.data_ptr(), R, N, K,
                               torch.cuda.current_stream(G.device).cuda_stream))
    return dW, db
