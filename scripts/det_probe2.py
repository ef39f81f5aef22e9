"""Which runs differ: sequences of deterministic-mode forwards at the headline shape, back to back (no host
synchronisation between runs) and with a synchronisation after every run, in a process that did / did not run the
default mode and a CUDA-graph phase before.  Usage: python scripts/det_probe2.py [prior] out.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tdanet_b200.look2hear as look2hear
from tdanet_b200 import _lib
from det_probe import timed

DEV = "cuda:0"


def seq(m, x, n, sync):
    outs = []
    with torch.no_grad():
        for _ in range(n):
            outs.append(m(x).clone())
            if sync:
                torch.cuda.synchronize()
    torch.cuda.synchronize()
    ref = outs[0]
    bad = [i for i, o in enumerate(outs) if not torch.equal(ref, o)]
    groups = []
    for o in outs:
        for gi, g in enumerate(groups):
            if torch.equal(g, o):
                break
        else:
            groups.append(o)
    return {"differ_from_run0": bad, "distinct_results": len(groups)}


def main():
    prior = len(sys.argv) > 2 and sys.argv[1] == "prior"
    out = sys.argv[-1]
    torch.manual_seed(0)
    m = look2hear.models.TDANetBest(out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=5,
                                    enc_kernel_size=4, num_sources=2, sample_rate=16000).eval().to(DEV)
    x = (torch.randn(64, 1, 32000, generator=torch.Generator().manual_seed(1)) * 0.1).to(DEV)
    rec = {"prior_default_and_graph_phase": prior}
    if prior:
        seq(m, x, 6, False)
        timed(m, x)
    _lib.set_deterministic(True)
    rec["back_to_back"] = seq(m, x, 40, False)
    rec["synchronised"] = seq(m, x, 40, True)
    rec["back_to_back_again"] = seq(m, x, 40, False)
    print(json.dumps(rec))
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
