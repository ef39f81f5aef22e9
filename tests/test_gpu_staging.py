"""The streaming kernels have two ways of staging their rows in shared memory (round 2): cp.async.bulk by one thread of
the CTA onto an mbarrier ring (512-channel models, the default) and per-thread cp.async rings (every other channel
count, and TDANET_BULK=0).  Both evaluate the same expressions in the same order, so the outputs of one model under
the two forms must agree to the run-to-run noise of the GlobLN statistics (double atomics).  The knob is read once per
process, hence the subprocesses.  Kernels: csrc/dwconv_impl.cuh (SbRing / SbRingN, la_stream_kernel, la_local_stats_kernel,
gstats_stream_kernel, dw5_pool_kernel); reference: look2hear/models/TDANet_best.py:266-292,342-380."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r})
import tdanet_b200.look2hear as look2hear
torch.manual_seed(0)
cls = getattr(look2hear.models, {cls!r})
m = cls(sample_rate=16000, out_channels=128, in_channels=512, num_blocks=3, upsampling_depth=5, enc_kernel_size=4,
        num_sources=2).eval().to("cuda:0")
m.gemm_mode = "fp32" if {act!r} == "fp32" else "tf32"    # bf16 storage needs a tensor-core GEMM mode
m.act_dtype = {act!r}
x = (torch.randn(5, 1, 24000, generator=torch.Generator().manual_seed(7)) * 0.1).to("cuda:0")
with torch.no_grad():
    y = m(x).float().cpu()
torch.save(y, {out!r})
"""


def _run(tmp_path, tag, env, cls="TDANetBest", act="fp32"):
    out = str(tmp_path / f"{tag}.pt")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT, cls=cls, act=act, out=out)], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


@pytest.mark.parametrize("cls,act", [("TDANetBest", "fp32"), ("TDANet", "fp32"), ("TDANetBest", "bf16")])
def test_bulk_and_per_thread_staging_agree(tmp_path, cls, act):
    ref = _run(tmp_path, "bulk", {"TDANET_BULK": "15"}, cls, act)
    assert torch.isfinite(ref).all() and ref.abs().max() > 0
    tol = 2e-5 if act == "fp32" else 2e-2     # bf16 storage: a statistic that moves by 1e-7 can flip a rounding
    # none; for the headline model also each kernel family alone
    for mask in (("0", "1", "2", "4", "8") if (cls, act) == ("TDANetBest", "fp32") else ("0",)):
        y = _run(tmp_path, f"mask{mask}", {"TDANET_BULK": mask}, cls, act)
        err = ((y - ref).abs().max() / ref.abs().max()).item()
        assert err < tol, (cls, act, mask, err)



def test_second_device_in_the_same_process():
    """The opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: a process that has run the
    512-channel kernels on cuda:0 must be able to run them on cuda:1 (csrc/common.cuh: PerDeviceOnce)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import tdanet_b200.look2hear as look2hear
    torch.manual_seed(0)
    kw = dict(sample_rate=16000, out_channels=128, in_channels=512, num_blocks=2, upsampling_depth=5, enc_kernel_size=4,
              num_sources=2)
    m0 = look2hear.models.TDANetBest(**kw).eval()
    sd = {k: v.clone() for k, v in m0.state_dict().items()}
    x = torch.randn(3, 1, 16000, generator=torch.Generator().manual_seed(3)) * 0.1
    ys = []
    for dev in ("cuda:0", "cuda:1"):
        m = look2hear.models.TDANetBest(**kw).eval()
        m.load_state_dict(sd)
        m = m.to(dev)
        m.gemm_mode = "tf32"
        with torch.no_grad():
            ys.append(m(x.to(dev)).float().cpu())
    assert torch.isfinite(ys[1]).all()
    assert ((ys[0] - ys[1]).abs().max() / ys[0].abs().max()).item() < 1e-3


@pytest.mark.parametrize("cls,variant", [("TDANetBest", "best"), ("TDANet", "fork")])
@pytest.mark.parametrize("enc_ms,T", [(4, 640), (4, 3333), (4, 12345), (4, 40001), (2, 9001)])
def test_bulk_staged_kernels_on_ragged_lengths(cls, variant, enc_ms, T):
    """512-channel models (the bulk-staged kernels) on lengths that are no multiple of the window, the hop or any tile:
    tensors of a few rows (every CTA an edge CTA), one interior chunk between two edges, the long case with many
    interior CTAs per item, the 2 ms encoder; batch 1 and 3, against the oracle in fp32."""
    import tdanet_b200.look2hear as look2hear
    from conftest import max_rel
    from oracle import tdanet_oracle as O
    kw = dict(out_channels=128, in_channels=512, num_blocks=2, upsampling_depth=5, enc_kernel_size=enc_ms, num_sources=2)
    torch.manual_seed(T)
    m = getattr(look2hear.models, cls)(sample_rate=16000, **kw).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    cfg = O.OracleConfig(variant=variant, sample_rate=16000, **kw)
    m = m.to("cuda:0")
    m.gemm_mode = "fp32"
    for B in (1, 3):
        x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(T + B)) * 0.1
        with torch.no_grad():
            ref = O.forward(sd, x, cfg)
            y = m(x.to("cuda:0")).cpu()
        assert y.shape == ref.shape == (B, 2, T)
        assert max_rel(y, ref) < 5e-5, (cls, enc_ms, T, B)
