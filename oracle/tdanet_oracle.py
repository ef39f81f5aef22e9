"""CPU oracle for the TDANet separation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) *restatement* of the reference
algorithm, written from the behaviour of the reference modules and working
directly on a ``state_dict`` with the reference's key names.  It is the checker
used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``tdanet_b200/``
imports it and it is never the thing that is shipped or measured as product.

Parity pinning: the reference has no golden vectors of its own,
so this restatement is pinned against outputs of the *unmodified reference*
(`/root/reference`, imported with the `oracle/ref_shim/timm` stub) generated in
the build container by ``oracle/make_golden.py`` and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks them on every run
and, when ``/root/reference`` is present, re-runs the live reference too.

Reference lines restated (all under /root/reference/look2hear/):

* ``forward``            models/TDANet_best.py:482-521, models/TDANet.py:869-909,
                         models/TDANet_mult_tes.py:540-579
* ``pad_input``          models/TDANet_best.py:465-479
* ``glob_ln``            models/TDANet_best.py:47-64 (hand written GlobLN) and
                         models/TDANet.py:59-60 (GroupNorm(1, C, eps=1e-8))
* ``uconv_block``        models/TDANet_best.py:342-380, models/TDANet.py:586-636,
                         models/TDANet_mult_tes.py:391-434
* ``global_attention``   models/TDANet_best.py:236-264, models/TDANet.py:372-406,
                         models/TDANet_mult_tes.py:254-285
* ``ffn``                models/TDANet_best.py:195-213 (``Mlp`` in the fork: TDANet.py:329-347)
* ``la``                 models/TDANet_best.py:266-292
* ``recurrent``          models/TDANet_best.py:383-399
* ``pairwise_neg_sdr``   losses/matrix.py:21-56
* ``pit_loss``           losses/pit_wrapper.py:29-67,106-131
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from itertools import permutations
from typing import Dict, Optional

import torch
import torch.nn.functional as F

EPS_GLN = 1e-8
EPS_LN = 1e-5


@dataclass
class OracleConfig:
    variant: str = "best"          # "best" | "fork" | "multres" | "origin" / "yang" (TDANetOrigin / TDANetYang:
                                   # GroupNorm norms, average-pool gather, additive injection, batch-axis attention)
    out_channels: int = 128
    in_channels: int = 512
    num_blocks: int = 16
    upsampling_depth: int = 5
    enc_kernel_size: int = 4       # milliseconds, like the reference kwarg
    num_sources: int = 2
    sample_rate: int = 16000
    kernels: int = 4               # multres only
    n_head: int = 8
    taps: Optional[Dict[str, torch.Tensor]] = field(default=None, repr=False)
    tap_block: int = 0             # per-stage activations are recorded for this block only
    tap_all: bool = False          # record every block instead, keys "<name>@<block>" (backward-pass tests)
    cur_block: Optional[int] = field(default=None, repr=False)
    # train-mode stochastic layers with EXPLICIT keep-masks (so a run can be compared element for element):
    # drop_masks[block] = {"att" [T'*heads, B, B], "ao" [B, T', C], "f1" [B, T', 2C], "f2" [B, T', C], "dp" [2, B]}
    # (bool / 0-1 tensors, channels-last like the CUDA workspace; missing keys = layer not applied).  Kept entries are
    # scaled by 1/(1-p): nn.Dropout(dropout) (TDANet_best.py:210,212,251), the dropout on the attention weights of
    # nn.MultiheadAttention(dropout) (:241), DropPath(drop_path) (:7-30,:262-263).
    drop_masks: Optional[list] = field(default=None, repr=False)
    dropout: float = 0.1
    drop_path: float = 0.1

    def mask(self, name, dtype=torch.float32):
        """Multiplier tensor (mask / keep_prob) of stochastic layer `name` in the current block, or None."""
        if self.drop_masks is None or self.cur_block is None:
            return None
        m = self.drop_masks[self.cur_block].get(name)
        if m is None:
            return None
        p = self.drop_path if name == "dp" else self.dropout
        return m.to(dtype) / (1.0 - p)

    @property
    def K(self) -> int:            # encoder window in samples
        return self.enc_kernel_size * self.sample_rate // 1000

    @property
    def S(self) -> int:            # encoder hop
        return self.K // 4

    @property
    def n_basis(self) -> int:
        return self.out_channels if self.variant == "multres" else self.K // 2 + 1


def _tap(cfg: OracleConfig, name: str, t: torch.Tensor) -> None:
    if cfg.taps is not None:
        if cfg.tap_all and cfg.cur_block is not None:
            name = f"{name}@{cfg.cur_block}"
        cfg.taps[name] = t.detach().clone()


# --------------------------------------------------------------------------- norms
def glob_ln(x, gamma, beta):
    """Global layer norm over (C, T) per batch item, biased variance, eps inside sqrt."""
    dims = tuple(range(1, x.ndim))
    mu = x.mean(dim=dims, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=dims, keepdim=True)
    xn = (x - mu) / torch.sqrt(var + EPS_GLN)
    shape = (1, -1) + (1,) * (x.ndim - 2)
    return xn * gamma.view(shape) + beta.view(shape)


def _gln_keys(variant):
    # hand-written GlobLN stores gamma/beta; GroupNorm(1, C) stores weight/bias
    return ("gamma", "beta") if variant == "best" else ("weight", "bias")


def _norm(sd, prefix, x, cfg):
    g, b = _gln_keys(cfg.variant)
    return glob_ln(x, sd[f"{prefix}.{g}"], sd[f"{prefix}.{b}"])


def prelu(x, w):
    return torch.where(x >= 0, x, w * x)


# --------------------------------------------------------------------------- pieces
def pad_input(wav, K, S):
    """Right pad to a whole number of windows, then K-S zeros on both sides."""
    n = wav.shape[1]
    rest = K - (S + n % K) % K
    if rest > 0:
        wav = F.pad(wav, (0, rest))
    wav = F.pad(wav, (K - S, K - S))
    return wav, rest


def conv_norm(sd, prefix, x, cfg, stride=1, groups=1):
    w = sd[f"{prefix}.conv.weight"]
    b = sd.get(f"{prefix}.conv.bias")
    y = F.conv1d(x, w, b, stride=stride, padding=(w.shape[-1] - 1) // 2, groups=groups)
    _tap(cfg, f"raw:{prefix}", y)
    return _norm(sd, f"{prefix}.norm", y, cfg)


def la(sd, prefix, x_l, x_g, cfg):
    """Local/global gated fusion: gLN(dw(x_l)) * sigmoid(up(gLN(dw(x_g)))) + up(gLN(dw(x_g)))."""
    C = x_l.shape[1]
    T = x_l.shape[-1]
    loc = conv_norm(sd, f"{prefix}.local_embedding", x_l, cfg, groups=C)
    act = conv_norm(sd, f"{prefix}.global_act", x_g, cfg, groups=C)
    emb = conv_norm(sd, f"{prefix}.global_embedding", x_g, cfg, groups=C)
    gate = F.interpolate(torch.sigmoid(act), size=T, mode="nearest")
    emb = F.interpolate(emb, size=T, mode="nearest")
    return loc * gate + emb


def mha_seq_first(x, w_in, b_in, w_out, b_out, n_head, cfg=None, tap=True):
    """nn.MultiheadAttention eval forward for input laid out (seq, batch, embed)."""
    S, N, E = x.shape
    d = E // n_head
    qkv = F.linear(x, w_in, b_in)
    if cfg is not None and tap:
        _tap(cfg, "ga.qkv", qkv)
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(S, N * n_head, d).transpose(0, 1) * (1.0 / math.sqrt(d))
    k = k.reshape(S, N * n_head, d).transpose(0, 1)
    v = v.reshape(S, N * n_head, d).transpose(0, 1)
    p = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    if cfg is not None and cfg.mask("att") is not None:
        p = p * cfg.mask("att", p.dtype)             # dropout on the attention weights, [N*heads, S, S]
    o = (p @ v).transpose(0, 1).reshape(S, N, E)
    if cfg is not None:
        _tap(cfg, "ga.attn_ctx", o if tap else o.transpose(0, 1))       # [B, T', C] either way
    return F.linear(o, w_out, b_out)


def global_attention(sd, prefix, x, cfg):
    """x: [B, C, T'] -> [B, C, T'] : x + MHA-block(x), then + FFN."""
    a = f"{prefix}.attn"
    C = x.shape[1]
    xt = x.transpose(1, 2)                                   # [B, T', C]
    h = F.layer_norm(xt, (C,), sd[f"{a}.attn_in_norm.weight"], sd[f"{a}.attn_in_norm.bias"], EPS_LN)
    h = h + sd[f"{a}.pos_enc.pe"][:, : h.shape[1]]           # PE indexed by time
    _tap(cfg, "ga.attn_in", h)
    w_in, b_in = sd[f"{a}.attn.in_proj_weight"], sd[f"{a}.attn.in_proj_bias"]
    w_out, b_out = sd[f"{a}.attn.out_proj.weight"], sd[f"{a}.attn.out_proj.bias"]
    if cfg.variant == "multres":
        # batch_first=True: sequence axis is time; correct residual
        o = mha_seq_first(h.transpose(0, 1), w_in, b_in, w_out, b_out, cfg.n_head, cfg, tap=False).transpose(0, 1)
        _tap(cfg, "ga.qkv", F.linear(h, w_in, b_in))
        m_ao = cfg.mask("ao", o.dtype)
        if m_ao is not None:
            o = o * m_ao                                   # output + self.dropout(attn_output)
        post = h + o
    else:
        # batch_first=False fed [B, T', C]: the *batch* axis is the sequence axis,
        # and the "residual" doubles the attention output (bug-compatible)
        o = mha_seq_first(h, w_in, b_in, w_out, b_out, cfg.n_head, cfg)
        m_ao = cfg.mask("ao", o.dtype)
        post = o + (o if m_ao is None else o * m_ao)                       # output + self.dropout(output)
        if m_ao is not None:
            o = post           # what the CUDA workspace keeps in train mode: out * (1 + mask / keep)
    _tap(cfg, "ga.attn_out", o)
    post = F.layer_norm(post, (C,), sd[f"{a}.norm.weight"], sd[f"{a}.norm.bias"], EPS_LN)
    dp = cfg.mask("dp", post.dtype)
    if dp is not None:
        post = post * dp[0].view(-1, 1, 1)                                 # self.drop_path(self.attn(x))
    x = x + post.transpose(1, 2)
    _tap(cfg, "ga.after_attn", x)
    # FFN
    m = f"{prefix}.mlp"
    y = conv_norm(sd, f"{m}.fc1", x, cfg)
    y = F.conv1d(y, sd[f"{m}.dwconv.weight"], sd[f"{m}.dwconv.bias"], padding=2, groups=y.shape[1])
    y = torch.relu(y)
    if cfg.mask("f1") is not None:
        y = y * cfg.mask("f1", y.dtype).transpose(1, 2)                    # FFN.drop after the activation
    _tap(cfg, "ga.ffn_dw", y)
    y = conv_norm(sd, f"{m}.fc2", y, cfg)
    if cfg.mask("f2") is not None:
        y = y * cfg.mask("f2", y.dtype).transpose(1, 2)                    # FFN.drop after fc2
    if dp is not None:
        y = y * dp[1].view(-1, 1, 1)                                       # self.drop_path(self.mlp(x))
    return x + y


def uconv_block(sd, prefix, x, cfg):
    depth = cfg.upsampling_depth
    C = cfg.in_channels
    residual = x
    p = f"{prefix}.proj_1x1"
    y = F.conv1d(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"])
    _tap(cfg, "proj.raw", y)
    y = prelu(_norm(sd, f"{p}.norm", y, cfg), sd[f"{p}.act.weight"])
    outs = [conv_norm(sd, f"{prefix}.spp_dw.0", y, cfg, stride=1, groups=C)]
    for k in range(1, depth):
        outs.append(conv_norm(sd, f"{prefix}.spp_dw.{k}", outs[-1], cfg, stride=2, groups=C))
    for k, o in enumerate(outs):
        _tap(cfg, f"spp.{k}", o)
    T_bot = outs[-1].shape[-1]
    if cfg.variant == "fork":
        # learned pooling: conv_pool[depth-1-k] = dw (k=2s+1, stride s) -> 1x1 -> gLN
        g = 0
        for k, o in enumerate(outs):
            q = f"{prefix}.conv_pool.{depth - 1 - k}"
            wd = sd[f"{q}.dw_conv.weight"]
            s = 2 ** (depth - 1 - k)                      # conv_pool[j] strides by 2^j
            z = F.conv1d(o, wd, sd[f"{q}.dw_conv.bias"], stride=s, padding=(wd.shape[-1] - 1) // 2, groups=C)
            _tap(cfg, f"pool.dw.{k}", z)
            z = F.conv1d(z, sd[f"{q}.pw_conv.weight"], sd[f"{q}.pw_conv.bias"])
            _tap(cfg, f"pool.pw.{k}", z)
            g = g + _norm(sd, f"{q}.norm", z, cfg)
    else:
        g = 0
        for o in outs:
            g = g + F.adaptive_avg_pool1d(o, T_bot)
    _tap(cfg, "ga.in", g)
    g = global_attention(sd, f"{prefix}.globalatt", g, cfg)
    _tap(cfg, "ga.out", g)
    fused = []
    for k in range(depth):
        if cfg.variant == "best":
            fused.append(la(sd, f"{prefix}.loc_glo_fus.{k}", outs[k], g, cfg))
        else:
            fused.append(F.interpolate(g, size=outs[k].shape[-1], mode="nearest") + outs[k])
        _tap(cfg, f"fused.{k}", fused[-1])
    expanded = None
    for i in range(depth - 2, -1, -1):
        # first step takes the *finer* neighbour as its "global" input (reference quirk)
        xg = fused[i - 1] if i == depth - 2 else expanded
        expanded = la(sd, f"{prefix}.last_layer.{i}", fused[i], xg, cfg)
        _tap(cfg, f"expanded.{i}", expanded)
    return F.conv1d(expanded, sd[f"{prefix}.res_conv.weight"], sd[f"{prefix}.res_conv.bias"]) + residual


def recurrent(sd, prefix, x, cfg):
    mixture = x
    wc, bc = sd[f"{prefix}.concat_block.0.weight"], sd[f"{prefix}.concat_block.0.bias"]
    ac = sd[f"{prefix}.concat_block.1.weight"]
    for i in range(cfg.num_blocks):
        if i > 0:
            x = prelu(F.conv1d(mixture + x, wc, bc, groups=x.shape[1]), ac)
        taps = cfg.taps
        cfg.cur_block = i
        if i != cfg.tap_block and not cfg.tap_all:
            cfg.taps = None                    # per-stage taps are recorded for one block only
        _tap(cfg, "block_in", x)
        x = uconv_block(sd, f"{prefix}.unet", x, cfg)
        cfg.taps = taps
        cfg.cur_block = None
        _tap(cfg, f"block.{i}", x)
    return x


def encode(sd, x, cfg):
    K, S = cfg.K, cfg.S
    if cfg.variant == "multres":
        embs = []
        for k in range(cfg.kernels):
            w = sd[f"encoder.conv_list.{k}.weight"]
            embs.append(F.conv1d(x, w, None, stride=S, padding=w.shape[-1] // 2))
        return torch.cat(embs, dim=1)
    return F.conv1d(x, sd["encoder.weight"], None, stride=S, padding=K // 2)


def forward(sd: Dict[str, torch.Tensor], wav: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """est_sources = model(wav); wav is [T], [B, T] or [B, 1, T]."""
    one_d = wav.ndim == 1
    if one_d:
        wav = wav.unsqueeze(0)
    if wav.ndim == 3:
        wav = wav.squeeze(1)
    K, S = cfg.K, cfg.S
    x, rest = pad_input(wav, K, S)
    s = encode(sd, x.unsqueeze(1), cfg)
    _tap(cfg, "enc", s)
    x = _norm(sd, "ln", s, cfg)
    if cfg.variant != "multres":
        x = F.conv1d(x, sd["bottleneck.weight"], sd["bottleneck.bias"])
    _tap(cfg, "bottleneck", x)
    x = recurrent(sd, "sm", x, cfg)
    x = prelu(x, sd["mask_net.0.weight"])
    x = F.conv1d(x, sd["mask_net.1.weight"], sd["mask_net.1.bias"])
    _tap(cfg, "mlogit", x)
    B = x.shape[0]
    x = torch.relu(x.view(B, cfg.num_sources, cfg.n_basis, -1)) * s.unsqueeze(1)
    _tap(cfg, "masked", x)
    y = F.conv_transpose1d(x.view(B, -1, x.shape[-1]), sd["decoder.weight"], None, stride=S, padding=K // 2)
    y = y[:, :, K - S: -(rest + K - S)].contiguous()
    return y.squeeze(0) if one_d else y


# --------------------------------------------------------------------------- loss
def pairwise_neg_sdr(ests, targets, sdr_type="snr", eps=1e-8):
    """[B, n_src, T] x2 -> [B, est, tgt] negative (SI-)SDR / SNR in dB."""
    assert sdr_type in ("snr", "sisdr", "sdsdr")
    targets = targets - targets.mean(dim=2, keepdim=True)
    ests = ests - ests.mean(dim=2, keepdim=True)
    t = targets.unsqueeze(1)          # [B, 1, tgt, T]
    e = ests.unsqueeze(2)             # [B, est, 1, T]
    if sdr_type in ("sisdr", "sdsdr"):
        dot = (e * t).sum(dim=3, keepdim=True)
        energy = (t ** 2).sum(dim=3, keepdim=True) + eps
        proj = dot * t / energy
    else:
        proj = t.expand(-1, e.shape[1], -1, -1)
    noise = e - t if sdr_type in ("sdsdr", "snr") else e - proj
    ratio = (proj ** 2).sum(dim=3) / ((noise ** 2).sum(dim=3) + eps)
    return -10.0 * torch.log10(ratio + eps)


def pit_loss(ests, targets, sdr_type="snr", threshold_byloss=True, return_ests=False):
    """PITLossWrapper(pairwise_neg_<sdr_type>, pit_from='pw_mtx') for n_src <= 3."""
    n_src = targets.shape[1]
    pw = pairwise_neg_sdr(ests, targets, sdr_type)
    perms = list(permutations(range(n_src)))
    # loss of permutation p: mean_j pw[b, est = p[j], tgt = j]
    loss_set = torch.stack([sum(pw[:, p[j], j] for j in range(n_src)) / n_src for p in perms], dim=1)
    min_loss, idx = loss_set.min(dim=1)
    kept = min_loss
    if threshold_byloss and (min_loss > -30).any():
        kept = min_loss[min_loss > -30]
    loss = kept.mean()
    if not return_ests:
        return loss
    perm_t = torch.tensor(perms, dtype=torch.long)[idx]                # [B, n_src]
    reordered = torch.gather(ests, 1, perm_t.unsqueeze(-1).expand(-1, -1, ests.shape[-1]))
    return loss, reordered


def si_snr_db(est, tgt, eps=1e-8):
    """Plain (non-PIT) SI-SNR in dB per [..., T] pair; used for the bf16 acceptance check."""
    est = est - est.mean(dim=-1, keepdim=True)
    tgt = tgt - tgt.mean(dim=-1, keepdim=True)
    proj = (est * tgt).sum(-1, keepdim=True) * tgt / ((tgt ** 2).sum(-1, keepdim=True) + eps)
    return 10.0 * torch.log10((proj ** 2).sum(-1) / (((est - proj) ** 2).sum(-1) + eps) + eps)


# --------------------------------------------------------------------------- long-form (CSS)
def css_segments(wav, seg_len, overlap):
    """Chunking of LibriCSSDataset (datas/libricssdatamodule.py:73-106): hop = int(seg_len*(1-overlap)),
    the last chunk is zero padded and ends the loop.  Returns ([n, seg_len] tensor, pad_len)."""
    hop = int(seg_len * (1 - overlap))
    n, start, pad_len, segs = wav.shape[-1], 0, 0, []
    while start < n:
        seg = wav[start:start + seg_len]
        if start + seg_len > n:
            pad_len = start + seg_len - n
            seg = torch.cat([seg, torch.zeros(pad_len, dtype=seg.dtype)])
            start += pad_len
        segs.append(seg)
        start += hop
    return torch.stack(segs), pad_len


def css_stitch(ests, overlap_len, pad_len, trim_like_reference=False):
    """Stitching of audio_test_css.py:108-134.  ests: [n_chunks, 2, seg_len] (each chunk separated alone).
    The permutation of every later chunk is aligned by cosine similarity of its head with the tail of the
    FIRST chunk (the reference never updates s*_t_minus_1).  `trim_like_reference` reproduces the
    `[:, :-pad_len]` slice literally (empty output when pad_len == 0)."""
    out1, out2 = ests[0, 0], ests[0, 1]
    p1, p2 = ests[0, 0, -overlap_len:], ests[0, 1, -overlap_len:]
    for k in range(1, ests.shape[0]):
        e1, e2 = ests[k, 0], ests[k, 1]
        c1 = F.cosine_similarity(p1, e1[:overlap_len], dim=0) + F.cosine_similarity(p2, e2[:overlap_len], dim=0)
        c2 = F.cosine_similarity(p1, e2[:overlap_len], dim=0) + F.cosine_similarity(p2, e1[:overlap_len], dim=0)
        if c1 > c2:
            out1, out2 = torch.cat([out1, e1[overlap_len:]]), torch.cat([out2, e2[overlap_len:]])
        else:
            out1, out2 = torch.cat([out1, e2[overlap_len:]]), torch.cat([out2, e1[overlap_len:]])
    out = torch.stack([out1, out2])
    if trim_like_reference or pad_len > 0:
        out = out[:, :-pad_len] if pad_len > 0 else out[:, :0]
    return out


def css_separate(sd, wav, cfg, segment, overlap):
    """audio_test_css.main for one recording: chunk, separate every chunk alone (B=1), stitch."""
    seg_len = int(segment * cfg.sample_rate)
    segs, pad_len = css_segments(wav, seg_len, overlap)
    ests = torch.stack([forward(sd, s, cfg) for s in segs])
    return css_stitch(ests, int(cfg.sample_rate * segment * overlap), pad_len)
