"""Long-form (continuous speech separation) runner: the semantics of `audio_test_css.py:99-136` with
`LibriCSSDataset`'s chunking (`datas/libricssdatamodule.py:73-106`), batched on the device.

The reference separates one chunk per model call (B = 1), so chunks never attend to each other; here
all chunks of all recordings go through ONE forward with `attn_group = 1`, which is numerically the same
thing, and the cosine-similarity stitch runs in two CUDA kernels (`tdanet_css_stitch`) with no per-chunk
device-to-host copies.
"""
import ctypes as C
from typing import List, Tuple

import torch

from ... import _lib


def css_segments(n_samples: int, seg_len: int, overlap: float) -> Tuple[List[int], int]:
    """Chunk starts and the zero padding of the last chunk (libricssdatamodule.py:73-100)."""
    hop = int(seg_len * (1 - overlap))
    if hop <= 0:
        raise ValueError("overlap must be < 1")
    starts, start, pad_len = [], 0, 0
    while start < n_samples:
        starts.append(start)
        if start + seg_len > n_samples:
            pad_len = start + seg_len - n_samples
            start += pad_len
        start += hop
    return starts, pad_len


def separate_long(model, wavs: torch.Tensor, segment: float = 2.0, overlap: float = 0.25,
                  max_chunks_per_call: int = 1024, trim_like_reference: bool = False):
    """wavs [n_streams, n_samples] (or [n_samples]) on the GPU -> (separated [n_streams, 2, n_out], swap flags).

    `trim_like_reference=True` reproduces the reference's `output[:, :-pad_len]` literally, i.e. an empty
    result when the recording length needs no padding (pad_len == 0); the default keeps the audio.
    """
    one = wavs.ndim == 1
    if one:
        wavs = wavs.unsqueeze(0)
    if not wavs.is_cuda:
        raise _lib.TdanetError("separate_long runs on CUDA tensors only (no CPU path)")
    sr = model.sample_rate()
    seg_len = int(segment * sr)
    overlap_len = int(sr * segment * overlap)
    n_streams, n = wavs.shape
    starts, pad_len = css_segments(n, seg_len, overlap)
    n_chunks = len(starts)
    padded = torch.nn.functional.pad(wavs.float(), (0, seg_len))
    idx = torch.tensor(starts, device=wavs.device).unsqueeze(1) + torch.arange(seg_len, device=wavs.device)
    chunks = padded[:, idx]                                    # [n_streams, n_chunks, seg_len]
    # positions past the recording must be zeros (they are: `padded`), like the reference's zero pad
    flat = chunks.reshape(n_streams * n_chunks, seg_len)
    ests = torch.empty(n_streams * n_chunks, 2, seg_len, device=wavs.device)
    old_group = model.attn_group
    model.attn_group = 1                                       # every chunk alone, as in the reference loop
    try:
        with torch.no_grad():
            for lo in range(0, flat.shape[0], max_chunks_per_call):
                ests[lo:lo + max_chunks_per_call] = model(flat[lo:lo + max_chunks_per_call])
    finally:
        model.attn_group = old_group
    total = seg_len + (n_chunks - 1) * (seg_len - overlap_len)
    out_len = total - pad_len if (pad_len > 0 or not trim_like_reference) else 0
    swap = torch.empty(n_streams, n_chunks, dtype=torch.int32, device=wavs.device)
    out = torch.empty(n_streams, 2, out_len, device=wavs.device)
    lib = _lib.load()
    with torch.cuda.device(wavs.device):
        _lib.check(lib.tdanet_css_stitch(ests.data_ptr(), n_streams, n_chunks, seg_len, overlap_len, out_len,
                                         swap.data_ptr(), out.data_ptr() if out_len else None,
                                         torch.cuda.current_stream(wavs.device).cuda_stream))
    return (out[0], swap[0]) if one else (out, swap)
