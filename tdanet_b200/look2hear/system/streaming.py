"""Pipelined separation of a sequence of host batches (evaluation over a dataset, `audio_test.py:85-110` does one
`model(mix)` + `.cpu()` per utterance): the host-to-device copy of batch i+1 and the device-to-host copy of batch
i-1 overlap the forward of batch i (three streams, double-buffered device inputs).  Results are identical to calling
`model(batch)` on every batch in turn - same batches, same attention groups - only the copies move off the critical
path.  No CPU path: the model must be on a CUDA device."""
from typing import Iterable, List, Optional, Sequence

import torch

from ... import _lib


def separate_pipelined(model, host_batches: Iterable[torch.Tensor],
                       host_outputs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """host_batches: CPU tensors [B, T] / [B, 1, T] (pinned memory makes the copies asynchronous).
    host_outputs: optional pre-allocated pinned CPU tensors [B, n_src, T], used round-robin (a slot is only rewritten
    after its previous copy has completed); otherwise a pinned tensor is allocated per batch.
    Returns the list of CPU result tensors; they are complete when the function returns."""
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise _lib.TdanetError("separate_pipelined runs a CUDA model only (no CPU path); call model.cuda() first")
    if getattr(model, "use_cuda_graph", False):
        return _separate_pipelined_graphed(model, host_batches, host_outputs, dev)
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    xin = [None, None]
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [None, None]          # forward of the batch that last used xin[slot] has read it
    ev_out = {}                     # host_outputs slot -> its last device-to-host copy
    outs: List[torch.Tensor] = []
    with torch.no_grad():
        for i, h in enumerate(host_batches):
            if h.is_cuda:
                raise _lib.TdanetError("separate_pipelined takes host (CPU) batches; call model(batch) for device tensors")
            slot = i & 1
            with torch.cuda.stream(s_in):
                if ev_free[slot] is not None:
                    s_in.wait_event(ev_free[slot])
                if xin[slot] is None or xin[slot].shape != h.shape:
                    xin[slot] = torch.empty(h.shape, dtype=torch.float32, device=dev)
                xin[slot].copy_(h, non_blocking=True)
                ev_in[slot].record(s_in)
            cur.wait_event(ev_in[slot])
            est = model(xin[slot])
            ev_free[slot] = torch.cuda.Event()
            ev_free[slot].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[slot])
                if host_outputs is not None:
                    k = i % len(host_outputs)
                    ho = host_outputs[k]
                    if k in ev_out:
                        ev_out[k].synchronize()      # the consumer of that slot's previous result is the host
                else:
                    k, ho = None, torch.empty(est.shape, dtype=est.dtype).pin_memory()
                ho.copy_(est, non_blocking=True)
                est.record_stream(s_out)
                if k is not None:
                    ev_out[k] = torch.cuda.Event()
                    ev_out[k].record(s_out)
            outs.append(ho)
    s_out.synchronize()
    return outs


_COPY_STREAMS = {}


def _copy_streams(dev):
    """One pair of copy streams per device, reused across calls."""
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _COPY_STREAMS[dev]


def _separate_pipelined_graphed(model, host_batches, host_outputs, dev):
    """The CUDA-graph form of the pipeline: two captured forwards with their own static input / output buffers
    alternate, the host batch is copied straight into the static input of its slot and the result straight out of the
    static output - no staging copies, no allocation in the steady state (the `model(xin).clone()` form paid two
    device-to-device copies per step on the copy engines the transfers need, and measured 0.5 ms per step above the
    device time on a good run, several ms on a bad one)."""
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _copy_streams(dev)
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [None, None]          # replay that last read static_in[slot] / wrote static_out[slot] is done
    ev_drained = [None, None]       # device-to-host copy that last read static_out[slot] is done
    ev_out = {}                     # host_outputs slot -> its last device-to-host copy
    outs: List[torch.Tensor] = []
    n_src = model.num_sources
    with torch.no_grad():
        for i, h in enumerate(host_batches):
            if h.is_cuda:
                raise _lib.TdanetError("separate_pipelined takes host (CPU) batches; call model(batch) for device tensors")
            h2 = h.squeeze(1) if h.ndim == 3 else h
            if h2.ndim != 2:
                raise _lib.TdanetError(f"separate_pipelined takes batches [B, T] or [B, 1, T], got {tuple(h.shape)}")
            B, T = h2.shape
            slot = i & 1
            g, sin, sout = model.graph_slot(B, T, slot)
            with torch.cuda.stream(s_in):
                if ev_free[slot] is not None:
                    s_in.wait_event(ev_free[slot])
                sin.copy_(h2, non_blocking=True)
                sin.record_stream(s_in)   # a larger batch later re-captures and drops these buffers mid-copy
                ev_in[slot].record(s_in)
            cur.wait_event(ev_in[slot])
            if ev_drained[slot] is not None:
                cur.wait_event(ev_drained[slot])     # the previous result of this slot has left the device
            g.replay()
            ev_free[slot] = torch.cuda.Event()
            ev_free[slot].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[slot])
                if host_outputs is not None:
                    k = i % len(host_outputs)
                    ho = host_outputs[k]
                    if k in ev_out:
                        ev_out[k].synchronize()      # the consumer of that slot's previous result is the host
                else:
                    k, ho = None, torch.empty(B, n_src, T, dtype=torch.float32).pin_memory()
                ho.copy_(sout, non_blocking=True)
                sout.record_stream(s_out)
                ev_drained[slot] = torch.cuda.Event()
                ev_drained[slot].record(s_out)
                if k is not None:
                    ev_out[k] = ev_drained[slot]
            outs.append(ho)
    s_out.synchronize()
    return outs
