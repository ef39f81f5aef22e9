"""Hunt for rare run-to-run differences with exact statistics (deterministic mode on): many back-to-back forwards
of a short model (no host synchronisation in the loop), per-item checksums of every named workspace tensor computed
on the device after each run, compared with the first run at the end.  Reports, per deviating run, the first tensor
in dataflow order that differs and which batch items of it differ.
Usage: python scripts/det_hunt.py --blocks 1 --runs 20000 [--mode tf32] [--out file.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tdanet_b200.look2hear as look2hear
from tdanet_b200 import _lib

DEV = "cuda:0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1)
    ap.add_argument("--runs", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    _lib.set_deterministic(True)
    torch.manual_seed(0)
    depth = 5
    m = look2hear.models.TDANetBest(out_channels=128, in_channels=512, num_blocks=a.blocks, upsampling_depth=depth,
                                    enc_kernel_size=4, num_sources=2, sample_rate=16000).eval().to(DEV)
    m.gemm_mode = a.mode
    B, T = a.batch, 32000
    x = (torch.randn(B, 1, T, generator=torch.Generator().manual_seed(1)) * 0.1).to(DEV)
    names = (["enc", "x0", "proj"] + [f"spp{k}" for k in range(depth)] + [f"pool_pw{k}" for k in range(depth)] +
             ["ga_in", "attn_in", "qkv", "attn_ctx", "attn_out", "ga_mid", "fc1", "ffn_dw", "fc2", "ga_out"] +
             [f"inj_coef{k}" for k in range(depth)] + ["fused_a", "fused_b"] +
             [f"expanded{k}" for k in range(depth - 2, -1, -1)] + ["block_out", "masked"])
    eng = m.engine
    with torch.no_grad():
        est = m(x)                                  # allocates the workspace
        views = [eng.workspace_tensor(n, B, T, DEV).view(B, -1).view(torch.int32) for n in names]
        cols = names + ["est"]
        sums = torch.zeros(a.runs, len(cols), B, dtype=torch.int64, device=DEV)
        out = torch.empty_like(est)
        w = m._weights()
        for r in range(a.runs):
            eng.forward(w, x.squeeze(1), 0, out=out)
            for j, v in enumerate(views):
                torch.sum(v, dim=1, dtype=torch.int64, out=sums[r, j])
            torch.sum(out.view(B, -1).view(torch.int32), dim=1, dtype=torch.int64, out=sums[r, len(views)])
        torch.cuda.synchronize()
    bad = (sums != sums[0:1]).any(dim=2)            # [runs, cols]
    events = []
    for r in bad.any(dim=1).nonzero().flatten().tolist():
        row = bad[r].nonzero().flatten().tolist()
        first = row[0]
        items = (sums[r, first] != sums[0, first]).nonzero().flatten().tolist()
        events.append({"run": r, "first_tensor": cols[first], "items_of_it": items[:16], "n_items": len(items),
                       "tensors": [cols[j] for j in row]})
    rec = {"blocks": a.blocks, "runs": a.runs, "mode": a.mode, "batch": B, "deviating_runs": len(events),
           "env": {k: v for k, v in os.environ.items() if k.startswith("TDANET_")}, "events": events[:20]}
    print(json.dumps(rec))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
