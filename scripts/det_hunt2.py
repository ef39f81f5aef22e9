"""Like det_hunt.py, with the location of a deviation: after every run (no host synchronisation) the suspects are
compared element by element with their copies from the first run, on the device; per run and tensor the script keeps
the number of differing elements, rows, channels, the first / last differing (item, row) and the largest difference.
The GlobLN statistics slots are compared as raw 64-bit words.
Usage: python scripts/det_hunt2.py --runs 24000 [--out file.json]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tdanet_b200.look2hear as look2hear
from tdanet_b200 import _lib

DEV = "cuda:0"
def compare(t, base, rec):
    """t, base [B, L, C]; rec: int64 [B + 1] device row: differing rows per item, then differing elements in total
    (a whole item = a statistic of it moved; <= one CTA's rows of one item = that CTA read or wrote something else)"""
    d = t != base
    torch.sum(d.any(dim=2), dim=1, out=rec[:-1])
    rec[-1] = d.sum()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=24000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    _lib.set_deterministic(True)
    torch.manual_seed(0)
    depth = 5
    m = look2hear.models.TDANetBest(out_channels=128, in_channels=512, num_blocks=1, upsampling_depth=depth,
                                    enc_kernel_size=4, num_sources=2, sample_rate=16000).eval().to(DEV)
    m.gemm_mode = a.mode
    B, T = a.batch, 32000
    x = (torch.randn(B, 1, T, generator=torch.Generator().manual_seed(1)) * 0.1).to(DEV)
    eng = m.engine
    lib = _lib.load()
    acts = (["proj"] + [f"spp{k}" for k in range(depth)] + [f"pool_pw{k}" for k in range(depth)] + ["ga_out"] +
            [f"inj_coef{k}" for k in range(depth)] + ["fused_a", "fused_b"] + [f"expanded{k}" for k in range(depth - 2, -1, -1)])
    stats = (["st_proj"] + [f"st_spp{k}" for k in range(depth)] + [f"st_la_l{k}" for k in range(depth - 1)] +
             [f"st_la_g{k}" for k in range(depth - 1)] + ["st_fc1", "st_fc2"])

    def raw_stat(name):
        off, dims = C.c_size_t(), (C.c_int64 * 3)()
        _lib.check(lib.tdanet_workspace_tensor(C.byref(eng.cfg), B, T, name.encode(), C.byref(off), C.byref(dims)))
        n = dims[0] * dims[1] * dims[2]
        return eng._ws[torch.device(DEV)][off.value: off.value + 8 * n].view(torch.int64).view(dims[0], 1, dims[1] * dims[2])

    with torch.no_grad():
        out = m(x).clone()
        views = {n: eng.workspace_tensor(n, B, T, DEV) for n in acts}
        views.update({n: raw_stat(n) for n in stats})
        cols = stats[:1] + acts[:1] + [v for k in range(depth) for v in (f"st_spp{k}", f"spp{k}", f"pool_pw{k}")]
        cols += ["st_fc1", "st_fc2", "ga_out"] + [f"inj_coef{k}" for k in range(depth)] + ["fused_a", "fused_b"]
        for k in range(depth - 2, -1, -1):
            cols += [f"st_la_l{k}", f"st_la_g{k}", f"expanded{k}"]
        assert set(cols) == set(acts + stats), sorted(set(acts + stats) ^ set(cols))
        base = {n: views[n].clone() for n in cols}
        rec = torch.zeros(a.runs, len(cols), B + 1, dtype=torch.int64, device=DEV)
        w = m._weights()
        xin = x.squeeze(1)
        for r in range(a.runs):
            eng.forward(w, xin, 0, out=out)
            for j, n in enumerate(cols):
                compare(views[n], base[n], rec[r, j])
        torch.cuda.synchronize()
    rec = rec.cpu()
    events = []
    for r in (rec[:, :, -1] != 0).any(dim=1).nonzero().flatten().tolist():
        ev = {"run": r, "tensors": []}
        for j, n in enumerate(cols):
            q = rec[r, j].tolist()
            if q[-1]:
                ev["tensors"].append({"name": n, "n_elements": q[-1], "rows_per_item": views[n].shape[1],
                                      "differing_rows_by_item": {str(i): v for i, v in enumerate(q[:-1]) if v}})
        events.append(ev)
    res = {"runs": a.runs, "mode": a.mode, "batch": B, "deviating_runs": len(events),
           "env": {k: v for k, v in os.environ.items() if k.startswith("TDANET_")},
           "events": [{"run": e["run"], "tensors": e["tensors"][:6]} for e in events[:12]]}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
