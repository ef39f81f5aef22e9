"""Run-to-run reproducibility and cost of the deterministic-statistics mode at BASELINE.json configs[1]
(TDANetBest 4 ms / 16 blocks / 64 x 2 s, TF32 GEMMs and fp32 GEMMs): max |difference| between repeated runs of the
same call with the mode off / on, and ms per CUDA-graph-replayed step of both.  Usage: python scripts/det_probe.py out.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tdanet_b200.look2hear as look2hear
from tdanet_b200 import _lib

DEV = "cuda:0"


def runs(m, x, n):
    outs = []
    with torch.no_grad():
        for _ in range(n):
            outs.append(m(x).clone())
    torch.cuda.synchronize()
    return outs


def timed(m, x, steps=20, warmup=5):
    m.use_cuda_graph = True
    with torch.no_grad():
        for _ in range(warmup):
            m(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    m.use_cuda_graph = False
    return e0.elapsed_time(e1) / steps


def main(out_path):
    torch.manual_seed(0)
    m = look2hear.models.TDANetBest(out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=5,
                                    enc_kernel_size=4, num_sources=2, sample_rate=16000).eval().to(DEV)
    x = (torch.randn(64, 1, 32000, generator=torch.Generator().manual_seed(1)) * 0.1).to(DEV)
    rec = {"workload": "TDANetBest 4 ms / 16 blocks / 64 x 2 s, seed-0 init", "runs_per_mode": 6}
    for mode in ("tf32", "fp32"):
        m.gemm_mode = mode
        for det in (False, True):
            _lib.set_deterministic(det)
            outs = runs(m, x, 6)
            scale = outs[0].abs().max().item()
            diffs = [(outs[0] - o).abs().max().item() / scale for o in outs[1:]]
            key = f"{mode}_{'deterministic' if det else 'default'}"
            rec[key] = {"max_rel_diff_between_runs": max(diffs), "runs_bit_identical": sum(torch.equal(outs[0], o) for o in outs[1:]),
                        "of": len(outs) - 1}
            if mode == "tf32":
                rec[key]["ms_per_step_graph"] = round(timed(m, x), 4)
                rec[key]["launches_per_step"] = None
                n0 = _lib.launch_count()
                with torch.no_grad():
                    m(x)
                rec[key]["launches_per_step"] = _lib.launch_count() - n0
            rec[key]["workspace_bytes"] = m.engine.workspace_bytes(64, 32000)
    _lib.set_deterministic(False)
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/det_probe.json")
