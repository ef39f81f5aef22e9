"""CPU emulation of the backward-pass CUDA sources (tests only; see tdanet_b200/csrc/emu.h).

Builds tdanet_b200/build_emu/libtdanet_emu.so with g++ from the *same* .cu files the CUDA library is built
from (-DTD_EMU), fills a training workspace with the activations of an oracle forward, runs
`tdanet_backward` on the host and returns the parameter gradients.  Used by tests/test_backward_emu.py to
compare against autograd of the oracle before the kernels reach a GPU.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import torch

from oracle import tdanet_oracle as O
from tdanet_b200 import _lib
from tdanet_b200.engine import SeparationEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tdanet_b200", "csrc")
OUT = os.path.join(ROOT, "tdanet_b200", "build_emu")
LIB = os.path.join(OUT, "libtdanet_emu.so")
SOURCES = ["backward.cu", "plan_abi.cu", "optim.cu"]


def build_emu() -> str:
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    stamp = os.path.join(OUT, "build.sha256")
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-x", "c++", "-DTD_EMU", "-DTDANET_BUILD",
           "-Wno-unknown-pragmas", "-Wno-attributes", "-I", CSRC, *srcs, "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    return LIB


_emu = None


def load_emu():
    global _emu
    if _emu is None:
        lib = C.CDLL(build_emu())
        fptr = C.c_void_p
        lib.tdanet_last_error.restype = C.c_char_p
        lib.tdanet_train_workspace_bytes.argtypes = [C.POINTER(_lib.Config), C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        lib.tdanet_train_workspace_tensor.argtypes = [C.POINTER(_lib.Config), C.c_int, C.c_int, C.c_char_p, C.c_int,
                                                      C.POINTER(C.c_size_t), C.POINTER(C.c_int64 * 3), C.POINTER(C.c_int32)]
        lib.tdanet_backward.argtypes = [C.POINTER(_lib.Config), C.POINTER(_lib.Weights), C.POINTER(_lib.Weights), fptr, fptr,
                                        C.c_int, C.c_int, fptr, C.c_size_t, fptr]
        _emu = lib
    return _emu


def _check(lib, code):
    if code != 0:
        raise RuntimeError(f"emu error {code}: {lib.tdanet_last_error().decode()}")


def make_engine(kw, sample_rate, gemm_mode="fp32", variant="best", act_dtype="fp32"):
    K = kw["enc_kernel_size"] * sample_rate // 1000
    multres = variant == "multres"
    return SeparationEngine(variant, kw["out_channels"], kw["in_channels"], kw["num_blocks"], kw["upsampling_depth"],
                            K, kw["out_channels"] if multres else K // 2 + 1, kw["num_sources"],
                            enc_convs=kw.get("kernels", 4) if multres else 1, gemm_mode=gemm_mode, act_dtype=act_dtype)


class Workspace:
    """Host training workspace addressed by the names of csrc/plan.h."""

    def __init__(self, lib, cfg, B, T):
        self.lib, self.cfg, self.B, self.T = lib, cfg, B, T
        n = C.c_size_t()
        _check(lib, lib.tdanet_train_workspace_bytes(C.byref(cfg), B, T, C.byref(n)))
        self.buf = np.zeros(n.value + 256, dtype=np.uint8)
        self.base = (-self.buf.ctypes.data) % 256   # 256-byte aligned start
        self.nbytes = n.value

    @property
    def ptr(self):
        return self.buf.ctypes.data + self.base

    def view(self, name, block=0):
        off, dims, es = C.c_size_t(), (C.c_int64 * 3)(), C.c_int32()
        _check(self.lib, self.lib.tdanet_train_workspace_tensor(C.byref(self.cfg), self.B, self.T, name.encode(), block,
                                                                C.byref(off), C.byref(dims), C.byref(es)))
        n = dims[0] * dims[1] * dims[2]
        dt = {1: np.uint8, 2: np.uint16, 4: np.float32, 8: np.float64}[es.value]   # 2: bf16 bits
        raw = self.buf[self.base + off.value: self.base + off.value + n * es.value]
        return raw.view(dt).reshape(dims[0], dims[1], dims[2])

    def put(self, name, t, block=0):
        v = self.view(name, block)
        if v.dtype == np.uint16:   # a large activation stored as bf16 (round to nearest even, like the forward kernels)
            a = t.detach().cpu().float().to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
        else:
            a = t.detach().cpu().numpy()
        assert a.shape == v.shape, (name, a.shape, v.shape)
        v[...] = a


def _cl(t):
    """[B, C, T] -> channels-last [B, T, C]"""
    return t.transpose(1, 2).contiguous()


def _stats(*raws):
    """per-item (sum, sum of squares) in double for each raw conv output -> [B, n, 2]"""
    out = []
    for r in raws:
        r = r.double().flatten(1)
        out.append(torch.stack([r.sum(1), (r * r).sum(1)], dim=1))
    return torch.stack(out, dim=1)


def random_drop_masks(B, Lb, C, n_head, num_blocks, dropout, drop_path, seed=11, time_axis=False):
    """Keep-masks of every stochastic layer of every block in the layout of the CUDA workspace / OracleConfig.drop_masks."""
    g = torch.Generator().manual_seed(seed)
    keep = lambda shape, p: (torch.rand(shape, generator=g) >= p).to(torch.uint8)
    out = []
    for _ in range(num_blocks):
        m = {}
        if dropout > 0:
            m.update(att=keep((B * n_head, Lb, Lb) if time_axis else (Lb * n_head, B, B), dropout), ao=keep((B, Lb, C), dropout),
                     f1=keep((B, Lb, 2 * C), dropout), f2=keep((B, Lb, C), dropout))
        if drop_path > 0:
            m["dp"] = keep((2, B), drop_path)
        out.append(m)
    return out


def fill_workspace(ws: Workspace, taps, kw, variant="best", drop_masks=None):
    """What tdanet_forward_train leaves in the workspace, written from the taps of an oracle forward."""
    depth, nb = kw["upsampling_depth"], kw["num_blocks"]
    if drop_masks is not None:
        for b, m in enumerate(drop_masks):
            for key, name in (("att", "m_att"), ("ao", "m_ao"), ("f1", "m_f1"), ("f2", "m_f2")):
                if key in m:
                    ws.put(name, m[key], b)
            if "dp" in m:
                ws.put("m_dp", m["dp"].unsqueeze(-1), b)
    u = "sm.unet"
    ws.put("enc", _cl(taps["enc"]))
    ws.put("st_enc", _stats(taps["enc"]))
    ws.put("x0", _cl(taps["bottleneck"]))
    ws.put("mlogit", _cl(taps["mlogit"]))
    B = taps["enc"].shape[0]
    ws.put("masked", _cl(taps["masked"].reshape(B, -1, taps["masked"].shape[-1])))
    partner = (depth - 3 + depth) % depth
    live = set(range(depth - 1)) | {partner}
    for b in range(nb):
        t = lambda name: taps[f"{name}@{b}"]
        if b > 0:
            ws.put("bin", _cl(t("block_in")), b)
        ws.put("y", _cl(taps[f"block.{b}"]), b)
        ws.put("proj", _cl(t("proj.raw")), b)
        ws.put("st_proj", _stats(t("proj.raw")), b)
        for k in range(depth):
            raw = t(f"raw:{u}.spp_dw.{k}")
            ws.put(f"spp{k}", _cl(raw), b)
            ws.put(f"st_spp{k}", _stats(raw), b)
            if k in live:
                ws.put(f"fused{k}", _cl(t(f"fused.{k}")), b)
                if variant == "best":
                    q = f"raw:{u}.loc_glo_fus.{k}"
                    ws.put(f"st_lgf{k}", _stats(t(f"{q}.local_embedding"), t(f"{q}.global_act"), t(f"{q}.global_embedding")), b)
            if variant == "fork":
                ws.put(f"pool_dw{k}", _cl(t(f"pool.dw.{k}")), b)
                ws.put(f"pool_pw{k}", _cl(t(f"pool.pw.{k}")), b)
                ws.put(f"st_pool{k}", _stats(t(f"pool.pw.{k}")), b)
        for i in range(depth - 1):
            ws.put(f"expanded{i}", _cl(t(f"expanded.{i}")), b)
            q = f"raw:{u}.last_layer.{i}"
            ws.put(f"st_la_l{i}", _stats(t(f"{q}.local_embedding")), b)
            ws.put(f"st_la_g{i}", _stats(t(f"{q}.global_act"), t(f"{q}.global_embedding")), b)
        ws.put("ga_in", _cl(t("ga.in")), b)
        ws.put("attn_in", t("ga.attn_in"), b)
        ws.put("qkv", t("ga.qkv"), b)
        ws.put("attn_ctx", t("ga.attn_ctx"), b)
        ws.put("attn_out", t("ga.attn_out"), b)
        ws.put("ga_mid", _cl(t("ga.after_attn")), b)
        m = f"{u}.globalatt.mlp"
        ws.put("fc1", _cl(t(f"raw:{m}.fc1")), b)
        ws.put("st_fc1", _stats(t(f"raw:{m}.fc1")), b)
        ws.put("ffn_dw", _cl(t("ga.ffn_dw")), b)
        ws.put("fc2", _cl(t(f"raw:{m}.fc2")), b)
        ws.put("st_fc2", _stats(t(f"raw:{m}.fc2")), b)
        ws.put("ga_out", _cl(t("ga.out")), b)


def emu_backward(sd, wav, d_est, kw, sample_rate, variant="best", dropout=0.0, drop_path=0.0, drop_masks=None,
                 act_dtype="fp32"):
    """Gradients of sum(est * d_est) w.r.t. every parameter, computed by the emulated CUDA backward pass.
    Returns (grads dict keyed like the state_dict, oracle output).  drop_masks: explicit keep-masks of a train-mode
    step (OracleConfig.drop_masks), written into the workspace where the forward pass would have drawn them."""
    lib = load_emu()
    # (bf16 storage needs a tensor-core gemm_mode in the configuration check; the emulated GEMMs are exact either way)
    eng = make_engine(kw, sample_rate, variant=variant, act_dtype=act_dtype, gemm_mode="tf32" if act_dtype == "bf16" else "fp32")
    eng.set_dropout(dropout, drop_path)
    okw = {k: v for k, v in kw.items() if k != "feat_len"}
    cfg = O.OracleConfig(variant=variant, sample_rate=sample_rate, taps={}, tap_all=True, drop_masks=drop_masks,
                         dropout=dropout, drop_path=drop_path, **okw)
    with torch.no_grad():
        est = O.forward(sd, wav, cfg)
    B, T = wav.shape[0], wav.shape[-1]
    ws = Workspace(lib, eng.cfg, B, T)
    fill_workspace(ws, cfg.taps, kw, variant, drop_masks)
    sd_c = {k: v.detach().contiguous().float() for k, v in sd.items()}
    grads = {k: torch.zeros_like(v) for k, v in sd_c.items() if not k.endswith("pos_enc.pe")}
    w = eng.pack(sd_c, _allow_host=True)
    gw = eng.pack(grads, _allow_host=True, optional=True)
    wav2 = wav.reshape(B, T).contiguous().float()
    d = d_est.contiguous().float()
    _check(lib, lib.tdanet_backward(C.byref(eng.cfg), C.byref(w), C.byref(gw), wav2.data_ptr(), d.data_ptr(), B, T,
                                    ws.ptr, ws.nbytes, None))
    return grads, est, ws
