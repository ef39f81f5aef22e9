"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python oracle/make_golden.py

The reference is imported from /root/reference with the `oracle/ref_shim` timm
stub on sys.path.  Three kinds of fixtures are written:

* ``<variant>_small.npz``  - a reduced-width model (weights + input + output +
  a few block-0 activations taken with forward hooks), small enough to commit;
* ``<variant>_full.npz``   - the BASELINE.json configuration, random-init under
  ``torch.manual_seed(0)``: per-tensor checksums of the reference state_dict
  (pins "identical random-init weights"), the seeded input recipe and the
  sub-sampled reference output;
* ``loss.npz``             - PairwiseNegSDR / PITLossWrapper known answers.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TDANET_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

import look2hear.models as RM          # noqa: E402  (the reference)
import look2hear.losses as RL          # noqa: E402
from oracle import tdanet_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CLASSES = {"best": "TDANetBest", "fork": "TDANet", "multres": "TDANetMultRes", "origin": "TDANetOrigin",
           "yang": "TDANetYang"}

SMALL = dict(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=5,
             enc_kernel_size=4, num_sources=2)
FULL = dict(out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=5,
            enc_kernel_size=4, num_sources=2)
# BASELINE config #1: configs/tdanet_debug.yml audionet_config, sample_rate 8000
DEBUG = dict(out_channels=128, in_channels=512, num_blocks=8, upsampling_depth=5,
             enc_kernel_size=4, num_sources=2, feat_len=3010, kernels=4)


def build(variant, kwargs, sr, seed=0):
    torch.manual_seed(seed)
    return getattr(RM, CLASSES[variant])(sample_rate=sr, **kwargs).eval()


def ocfg(variant, kwargs, sr):
    kw = {k: v for k, v in kwargs.items() if k != "feat_len"}
    return O.OracleConfig(variant=variant, sample_rate=sr, **kw)


def trimmed_sd(model, t_bot):
    sd = {}
    for k, v in model.state_dict().items():
        if k.endswith("pos_enc.pe"):
            v = v[:, :t_bot]          # the buffer is 10000 rows; only [:T'] is ever read
        sd[k] = v.clone()
    return sd


def checksums(model):
    keys, s1, s2 = [], [], []
    for k, v in model.state_dict().items():
        keys.append(k)
        s1.append(v.double().sum().item())
        s2.append((v.double() ** 2).sum().item())
    return np.array(keys), np.array(s1), np.array(s2)


def hook_taps(model, names):
    taps, handles = {}, []
    mods = dict(model.named_modules())
    for n in names:
        def fn(_m, _i, o, n=n):
            taps.setdefault(n, o.detach().clone())      # first call = block 0
        handles.append(mods[n].register_forward_hook(fn))
    return taps, handles


def main():
    os.makedirs(OUT, exist_ok=True)
    report = []
    only = [a.split("=", 1)[1].split(",") for a in sys.argv[1:] if a.startswith("--only=")]
    for variant in (only[0] if only else CLASSES):
        sr = 8000 if variant == "multres" else 16000
        kw = dict(SMALL)
        if variant == "multres":
            kw.update(feat_len=3010, kernels=4)
        # ---------------- small model: everything stored
        m = build(variant, kw, sr, seed=7)
        with torch.no_grad():     # make the affine parameters non-trivial
            g = torch.Generator().manual_seed(11)
            for k, p in m.named_parameters():
                if p.ndim == 1:
                    p.add_(0.2 * torch.randn(p.shape, generator=g))
        T = 3000
        x = torch.randn(3, 1, T, generator=torch.Generator().manual_seed(1234)) * 0.1
        taps, hs = hook_taps(m, ["sm.unet.proj_1x1", "sm.unet.spp_dw.4", "sm.unet.globalatt",
                                 "sm.unet.last_layer.3", "sm.unet.last_layer.0", "sm.unet"])
        with torch.no_grad():
            y = m(x)
        for h in hs:
            h.remove()
        t_bot = taps["sm.unet.spp_dw.4"].shape[-1]
        sd = trimmed_sd(m, t_bot)
        cfg = ocfg(variant, kw, sr)
        with torch.no_grad():
            yo = O.forward(sd, x, cfg)
            yo64 = O.forward({k: v.double() for k, v in sd.items()}, x.double(), cfg)
        d = (y - yo).abs().max().item() / y.abs().max().item()
        d64 = (y.double() - yo64).abs().max().item() / y.abs().max().item()
        report.append(f"{variant}_small: oracle vs reference max-rel {d:.3e}; fp64 oracle vs reference {d64:.3e}")
        arrs = {"x": x.numpy(), "y": y.numpy()}
        arrs.update({"sd/" + k: v.numpy() for k, v in sd.items()})
        arrs.update({"tap/" + k: v.numpy() for k, v in taps.items()})
        arrs["kwargs"] = np.array(repr({k: v for k, v in kw.items()}))
        arrs["sample_rate"] = np.array(sr)
        np.savez_compressed(os.path.join(OUT, f"{variant}_small.npz"), **arrs)

        # ---------------- full model: seeded init, checksums + sub-sampled output
        kwf = dict(DEBUG) if variant == "multres" else dict(FULL)
        m = build(variant, kwf, sr, seed=0)
        B = 1 if variant == "multres" else 2
        x = torch.randn(B, 1, 32000, generator=torch.Generator().manual_seed(1234)) * 0.1
        with torch.no_grad():
            y = m(x)
            sd = {k: v for k, v in m.state_dict().items()}
            yo = O.forward(sd, x, ocfg(variant, kwf, sr))
        d = (y - yo).abs().max().item() / y.abs().max().item()
        report.append(f"{variant}_full: oracle vs reference max-rel {d:.3e} (B={B}, T=32000)")
        keys, s1, s2 = checksums(m)
        np.savez_compressed(os.path.join(OUT, f"{variant}_full.npz"),
                            keys=keys, sum=s1, sumsq=s2, y_sub=y[:, :, ::16].numpy(),
                            y_absmax=np.array(y.abs().max().item()),
                            kwargs=np.array(repr(kwf)), sample_rate=np.array(sr),
                            batch=np.array(B), input_seed=np.array(1234), init_seed=np.array(0))

    if only:   # model fixtures of the named variants only; the loss fixtures and the report stay as they are
        with open(os.path.join(OUT, "REPORT.txt"), "a") as f:
            f.write("\n".join(report) + "\n")
        print("\n".join(report))
        return
    # ---------------- loss known answers
    g = torch.Generator().manual_seed(99)
    tgt = torch.randn(6, 2, 2000, generator=g) * 0.1
    est = tgt.flip(1) + 0.05 * torch.randn(6, 2, 2000, generator=g)        # swapped order
    est[1] = tgt[1] + 1e-5 * torch.randn(2, 2000, generator=g)             # below the -30 dB threshold
    est[2] = torch.randn(2, 2000, generator=g) * 0.1                       # unrelated
    est = est + 0.01                                                      # non-zero mean
    arrs = {"est": est.numpy(), "tgt": tgt.numpy()}
    for name in ("snr", "sisdr", "sdsdr"):
        fn = getattr(RL, f"pairwise_neg_{name}")
        arrs[f"pw_{name}"] = fn(est, tgt).numpy()
        for thr in (True, False):
            w = RL.PITLossWrapper(fn, pit_from="pw_mtx", threshold_byloss=thr)
            loss, reo = w(est, tgt, return_ests=True)
            arrs[f"pit_{name}_{int(thr)}"] = loss.numpy()
            # which estimate was matched to target 0 (the reordering for n_src = 2)
            arrs[f"perm0_{name}_{int(thr)}"] = (reo[:, 0] == est[:, 1]).all(dim=-1).long().numpy()
            lo, ro = O.pit_loss(est, tgt, name, thr, return_ests=True)
            report.append(f"loss {name} thr={thr}: ref {loss.item():.6f} oracle {lo.item():.6f} "
                          f"reorder-equal {bool((ro == reo).all())}")
    # every item below threshold -> nothing is dropped
    est2 = tgt + 1e-6 * torch.randn(6, 2, 2000, generator=g)
    w = RL.PITLossWrapper(RL.pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=True)
    arrs["est_all_below"] = est2.numpy()
    arrs["pit_snr_all_below"] = w(est2, tgt).numpy()
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **arrs)

    with open(os.path.join(OUT, "REPORT.txt"), "w") as f:
        f.write("generated by oracle/make_golden.py from the unmodified reference\n")
        f.write(f"torch {torch.__version__}\n")
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
