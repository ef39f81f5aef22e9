"""Eager PyTorch on this GPU, stage by stage, at the shapes of one UConvBlock of the headline configuration
(TDANetBest 4 ms, B = 64: C = 512, c = 128, L = 2010 .. 126) - the per-kernel bar SURVEY.md section 2a sets ("stock
PyTorch-on-B200 for the same module").  Each stage is the reference module's op sequence (TDANet_best.py line ranges in
the labels) as stock ATen / cuDNN / cuBLAS calls, TF32 allowed; CUDA events, 20 repetitions after 5 warm-ups.
Prints a JSON table; DESIGN.md puts our kernels' event-timed figures of bench.py next to it."""
import json, sys
import torch
import torch.nn.functional as F

dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
B, C, c = 64, 512, 128
L = [2010, 1005, 503, 252, 126]
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)


def gln(x, gamma, beta):   # TDANet_best.py:47-64 (the reference's hand-written GlobLN: ~10 elementwise / reduction ops)
    dims = list(range(1, x.dim()))
    mean = x.mean(dim=dims, keepdim=True)
    var = torch.pow(x - mean, 2).mean(dim=dims, keepdim=True)
    return gamma * ((x - mean) / (var + 1e-8).sqrt()) + beta


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
with torch.no_grad():
    x = rn(B, c, L[0])
    gam, bet = rn(1, C, 1), rn(1, C, 1)
    w_proj, b_proj, slope = rn(C, c, 1) / c ** 0.5, rn(C), torch.tensor([0.25], device=dev)
    out["proj_1x1 (conv 1x1 + GlobLN + PReLU, :304,349)"] = timeit(lambda: F.prelu(gln(F.conv1d(x, w_proj, b_proj), gam, bet), slope))
    # spp_dw chain: depthwise k5 (stride 1, then 2 x4) + GlobLN (:179-192, 350-356)
    wd = [rn(C, 1, 5) for _ in range(5)]
    bd = [rn(C) for _ in range(5)]
    h0 = rn(B, C, L[0])

    def spp():
        o = [gln(F.conv1d(h0, wd[0], bd[0], stride=1, padding=2, groups=C), gam, bet)]
        for k in range(1, 5):
            o.append(gln(F.conv1d(o[-1], wd[k], bd[k], stride=2, padding=2, groups=C), gam, bet))
        return o
    out["spp_dw[0..4] (5 x depthwise k5 + GlobLN, :350-356)"] = timeit(spp)
    outs = spp()
    out["bottom gather (sum of adaptive_avg_pool1d, :358-364)"] = timeit(lambda: sum(F.adaptive_avg_pool1d(o, L[4]) for o in outs))
    # GA block at the bottom scale: LN + PE, MHA over the batch axis, LN(2a), FFN (:195-264)
    mha = torch.nn.MultiheadAttention(C, 8, 0.1).to(dev).eval()
    ln1, ln2 = torch.nn.LayerNorm(C).to(dev), torch.nn.LayerNorm(C).to(dev)
    pe = rn(1, L[4], C)
    w1, w2, wdw, bdw = rn(2 * C, C, 1) / C ** 0.5, rn(C, 2 * C, 1) / (2 * C) ** 0.5, rn(2 * C, 1, 5), rn(2 * C)
    g2, b2 = rn(1, 2 * C, 1), rn(1, 2 * C, 1)
    xg = rn(B, C, L[4])

    def ga():
        t = xg.transpose(1, 2)
        t = ln1(t) + pe
        a = mha(t, t, t)[0]
        a = ln2(a + a).transpose(1, 2)
        y = xg + a
        f = gln(F.conv1d(y, w1), g2, b2)
        f = F.relu(F.conv1d(f, wdw, bdw, padding=2, groups=2 * C))
        f = gln(F.conv1d(f, w2), gam, bet)
        return y + f
    out["GlobalAttention block (LN+PE, MHA over the batch axis, LN(2a), FFN, :195-264)"] = timeit(ga)
    # one LA (k = 5) at the finest scale: 3 depthwise convs + 3 GlobLN + 2 nearest interpolations + gate (:266-292)
    wl, wa, we = rn(C, 1, 5), rn(C, 1, 5), rn(C, 1, 5)
    xl, xgl = rn(B, C, L[0]), rn(B, C, L[1])

    def la(xl=xl, xg=xgl, ks=5):
        p = (ks - 1) // 2
        le = gln(F.conv1d(xl, wl[..., :ks], padding=p, groups=C), gam, bet)
        ga_ = gln(F.conv1d(xg, wa[..., :ks], padding=p, groups=C), gam, bet)
        ge = gln(F.conv1d(xg, we[..., :ks], padding=p, groups=C), gam, bet)
        return le * torch.sigmoid(F.interpolate(ga_, size=xl.shape[-1], mode="nearest")) + F.interpolate(ge, size=xl.shape[-1], mode="nearest")
    out["LA last_layer[0] (2010 <- 1005: 3 dw k5 + 3 GlobLN + interpolate + gate, :266-292)"] = timeit(la)
    gf = rn(B, C, L[4])
    out["loc_glo_fus[0] (LA k = 1 on the finest scale, 2010 <- 126, :329-331,367-371)"] = timeit(lambda: la(xl, gf, 1))
    e0 = rn(B, C, L[0])
    w_res, b_res = rn(c, C, 1) / C ** 0.5, rn(c)
    out["res_conv + residual (:380)"] = timeit(lambda: F.conv1d(e0, w_res, b_res) + x)
print(json.dumps({k: round(v, 3) for k, v in out.items()}, indent=1))
