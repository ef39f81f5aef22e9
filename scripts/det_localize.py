"""Bisecting aid on top of the deterministic-statistics mode: repeats one forward (TDANetBest 4 ms, 64 x 2 s, TF32
GEMMs unless --mode says otherwise) with the mode on, checksums every named workspace tensor after each run and
reports, in dataflow order, which tensors ever differ from the first run, in how many runs, and where.
Usage: python scripts/det_localize.py --blocks 1 --runs 400 [--mode tf32] [--out file.json]"""
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
    ap.add_argument("--runs", type=int, default=400)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--act", default="fp32")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    _lib.set_deterministic(True)
    torch.manual_seed(0)
    depth = 5
    m = look2hear.models.TDANetBest(out_channels=128, in_channels=512, num_blocks=a.blocks, upsampling_depth=depth,
                                    enc_kernel_size=4, num_sources=2, sample_rate=16000).eval().to(DEV)
    m.gemm_mode, m.act_dtype = a.mode, a.act
    B, T = a.batch, 32000
    x = (torch.randn(B, 1, T, generator=torch.Generator().manual_seed(1)) * 0.1).to(DEV)
    names = (["enc", "x0", "proj"] + [f"spp{k}" for k in range(depth)] + [f"pool_pw{k}" for k in range(depth)] +
             ["ga_in", "attn_in", "qkv", "attn_ctx", "attn_out", "ga_mid", "fc1", "ffn_dw", "fc2", "ga_out"] +
             [f"inj_coef{k}" for k in range(depth)] + ["fused_a", "fused_b"] +
             [f"expanded{k}" for k in range(depth - 2, -1, -1)] + ["block_out", "masked"])
    eng = m.engine

    def snapshot():
        return {n: eng.workspace_tensor(n, B, T, DEV) for n in names}

    def checksums(ts, est):
        c = {n: int(t.view(torch.int32).sum(dtype=torch.int64).item()) for n, t in ts.items()}
        c["est"] = int(est.view(torch.int32).sum(dtype=torch.int64).item())
        return c

    with torch.no_grad():
        est0 = m(x).clone()
        base = {n: t.clone() for n, t in snapshot().items()}
        base["est"] = est0
        c0 = checksums(snapshot(), est0)
        hits = {}
        runs_differ = 0
        for r in range(1, a.runs + 1):
            est = m(x)
            ts = snapshot()
            c = checksums(ts, est)
            bad = [n for n in names + ["est"] if c[n] != c0[n]]
            if bad:
                runs_differ += 1
            for n in bad:
                h = hits.setdefault(n, {"runs": 0, "first_run": r})
                h["runs"] += 1
                if "where" not in h:
                    t = est if n == "est" else ts[n]
                    d = (t != base[n])
                    idx = d.nonzero()
                    h["n_elements"] = int(d.sum().item())
                    h["max_abs"] = float((t - base[n]).abs().max().item())
                    h["scale"] = float(base[n].abs().max().item())
                    h["where"] = {"first": idx[0].tolist(), "last": idx[-1].tolist(),
                                  "items": sorted(set(idx[:, 0].tolist()))[:8],
                                  "rows": [int(idx[:, 1].min()), int(idx[:, 1].max())] if idx.shape[1] > 1 else None}
    order = [n for n in names + ["est"] if n in hits]
    rec = {"blocks": a.blocks, "runs": a.runs, "mode": a.mode, "act": a.act, "batch": B, "runs_that_differ": runs_differ,
           "env": {k: v for k, v in os.environ.items() if k.startswith("TDANET_")},
           "tensors_that_differ_in_dataflow_order": [{"name": n, **hits[n]} for n in order]}
    print(json.dumps(rec))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
