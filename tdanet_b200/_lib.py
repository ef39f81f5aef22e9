"""ctypes binding of include/tdanet_b200.h (the C-ABI shared library built from csrc/).

There is no CPU path: if the library is missing or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TDANET_LIB") or os.path.join(_HERE, "lib", "libtdanet_b200.so")   # TDANET_LIB: experiment builds

MAX_DEPTH = 8
MAX_ENC = 4
VARIANTS = {"best": 0, "fork": 1, "multres": 2, "origin": 3, "yang": 3}
GEMM_MODES = {"fp32": 0, "tf32": 1, "tf32x3": 2}
ACT_DTYPES = {"fp32": 0, "bf16": 1}
SDR_TYPES = {"snr": 0, "sisdr": 1, "sdsdr": 2}

fptr = C.c_void_p  # device pointers travel as integers


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "variant", "out_channels", "in_channels", "num_blocks", "depth", "enc_kernel", "enc_stride",
        "n_basis", "num_sources", "enc_convs", "n_head", "gemm_mode", "attn_group", "act_dtype")] + [
        ("dropout", C.c_float), ("drop_path", C.c_float)]   # training-mode only (tdanet_forward_train_rng / tdanet_backward)


class ConvNorm(C.Structure):
    _fields_ = [("w", fptr), ("b", fptr), ("gamma", fptr), ("beta", fptr)]


class LA(C.Structure):
    _fields_ = [("local_embedding", ConvNorm), ("global_embedding", ConvNorm), ("global_act", ConvNorm)]


class SepConvNorm(C.Structure):
    _fields_ = [("dw_w", fptr), ("dw_b", fptr), ("pw_w", fptr), ("pw_b", fptr), ("gamma", fptr), ("beta", fptr)]


class Weights(C.Structure):
    _fields_ = [
        ("enc_w", fptr * MAX_ENC),
        ("ln_gamma", fptr), ("ln_beta", fptr),
        ("bottleneck_w", fptr), ("bottleneck_b", fptr),
        ("proj", ConvNorm), ("proj_prelu", fptr),
        ("spp_dw", ConvNorm * MAX_DEPTH),
        ("loc_glo_fus", LA * MAX_DEPTH),
        ("conv_pool", SepConvNorm * MAX_DEPTH),
        ("res_w", fptr), ("res_b", fptr),
        ("pe", fptr),
        ("ln1_w", fptr), ("ln1_b", fptr),
        ("in_proj_w", fptr), ("in_proj_b", fptr),
        ("out_proj_w", fptr), ("out_proj_b", fptr),
        ("ln2_w", fptr), ("ln2_b", fptr),
        ("fc1", ConvNorm), ("ffn_dw_w", fptr), ("ffn_dw_b", fptr), ("fc2", ConvNorm),
        ("last_layer", LA * MAX_DEPTH),
        ("concat_w", fptr), ("concat_b", fptr), ("concat_prelu", fptr),
        ("mask_prelu", fptr), ("mask_w", fptr), ("mask_b", fptr),
        ("dec_w", fptr),
        ("pe_rows", C.c_int32), ("reserved", C.c_int32),
    ]


class TdanetError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

_SIGNATURES = {
    "tdanet_abi_version": (C.c_int, []),
    "tdanet_abi_sizes": (C.c_int, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "tdanet_last_error": (C.c_char_p, []),
    "tdanet_launch_count": (C.c_uint64, []),
    "tdanet_device_supported": (C.c_int, [C.c_int]),
    "tdanet_set_deterministic": (C.c_int, [C.c_int]),
    "tdanet_get_deterministic": (C.c_int, []),
    "tdanet_profile_enable": (C.c_int, [C.c_int]),
    "tdanet_profile_dump": (C.c_int, [C.c_char_p, C.c_size_t]),
    "tdanet_workspace_bytes": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "tdanet_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), fptr, C.c_int, C.c_int, fptr, fptr,
                                 C.c_size_t, fptr]),
    "tdanet_workspace_tensor": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.c_char_p,
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_int64 * 3)]),
    "tdanet_latent_lengths": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_int32 * MAX_DEPTH),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tdanet_train_workspace_bytes": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "tdanet_forward_train": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), fptr, C.c_int, C.c_int, fptr, fptr,
                                       C.c_size_t, fptr]),
    "tdanet_forward_train_rng": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), fptr, C.c_int, C.c_int, fptr, fptr,
                                           C.c_size_t, fptr, fptr]),
    "tdanet_backward": (C.c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(Weights), fptr, fptr, C.c_int,
                                  C.c_int, fptr, C.c_size_t, fptr]),
    "tdanet_train_workspace_tensor": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.c_char_p, C.c_int,
                                                C.POINTER(C.c_size_t), C.POINTER(C.c_int64 * 3), C.POINTER(C.c_int32)]),
    "tdanet_wgrad": (C.c_int, [C.c_int, fptr, fptr, fptr, fptr, C.c_int, C.c_int, C.c_int, fptr]),
    "tdanet_grad_sqnorm": (C.c_int, [fptr, C.c_size_t, fptr, fptr]),
    "tdanet_adam_step": (C.c_int, [fptr, fptr, fptr, fptr, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_float, C.c_float, fptr, fptr, fptr]),
    "tdanet_gemm": (C.c_int, [C.c_int, fptr, fptr, fptr, fptr, C.c_int, C.c_int, C.c_int, C.c_int, fptr, fptr,
                              C.c_size_t, fptr]),
    "tdanet_gemm_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "tdanet_css_stitch": (C.c_int, [fptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fptr, fptr, fptr]),
    "tdanet_pit_loss_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "tdanet_pit_loss": (C.c_int, [fptr, fptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fptr, fptr, fptr, fptr,
                                  fptr, C.c_size_t, fptr]),
}
EXPORTS = tuple(_SIGNATURES)


def load():
    """dlopen the library (once), bind signatures and verify the struct mirrors."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise TdanetError(
                f"{LIB_PATH} is missing: build it with `python -m tdanet_b200._build` "
                "(there is no CPU or PyTorch fallback for the separation path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        cb, wb = C.c_size_t(), C.c_size_t()
        lib.tdanet_abi_sizes(C.byref(cb), C.byref(wb))
        if cb.value != C.sizeof(Config) or wb.value != C.sizeof(Weights):
            raise TdanetError(
                f"ABI mismatch: library structs are {cb.value}/{wb.value} bytes, "
                f"ctypes mirrors are {C.sizeof(Config)}/{C.sizeof(Weights)}")
        _lib = lib
        return lib


def check(code: int) -> None:
    if code != 0:
        msg = load().tdanet_last_error().decode(errors="replace")
        raise TdanetError(f"tdanet_b200 error {code}: {msg}")


def set_deterministic(on: bool) -> None:
    """Process-wide: exact (order-independent) accumulation of the GlobLN statistics of the inference forward, so
    that two runs of the same call return the same bits (include/tdanet_b200.h: tdanet_set_deterministic)."""
    check(load().tdanet_set_deterministic(int(bool(on))))


def deterministic() -> bool:
    return bool(load().tdanet_get_deterministic())


def launch_count() -> int:
    return int(load().tdanet_launch_count())


def profile_enable(on: bool) -> None:
    check(load().tdanet_profile_enable(int(on)))


def profile_dump():
    """[{kernel, launches, ms}] for every launch recorded since the last dump (synchronises)."""
    import json
    buf = C.create_string_buffer(1 << 16)
    n = load().tdanet_profile_dump(buf, len(buf))
    if n < 0:
        check(n)
    return json.loads(buf.value.decode())
