#!/usr/bin/env python
"""Headline benchmark: separated audio-seconds per second of TDANet inference on B200.

  python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A "step" is one forward of the hot path over one batch of synthetic 2 s / 16 kHz mixtures
(BASELINE.json configs[1]: TDANet 4 ms encoder, 16 blocks, batch 64 per GPU).  Rank 0 prints ONE
JSON line.  `value` is device-timed with the inputs resident in HBM; `e2e` goes through the public
`model(mix)` call with pinned host buffers (H2D of the mixtures and D2H of the separated sources inside
the timed region).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR = 16000
N_SAMPLES = 32000
METRIC = "separated_audio_seconds_per_second"
UNIT = "audio-s/s"
CLASSES = {"best": "TDANetBest", "fork": "TDANet"}


def model_kwargs(enc_ms):
    return dict(out_channels=128, in_channels=512, num_blocks=16, upsampling_depth=5, enc_kernel_size=enc_ms,
                num_sources=2)


def workload_name(variant, enc_ms, batch):
    return (f"{CLASSES[variant]} {enc_ms} ms/16 blocks, inference, {batch} x 2 s @16 kHz per GPU "
            f"(BASELINE.json configs[{1 if enc_ms == 4 else 2}]), seed-0 init")


def common_config(variant, enc_ms, batch):
    """The `config` object both arms (ours / reference) print verbatim, so that the driver's same-config check compares
    like with like; arm-specific knobs go to `impl_config`."""
    return {"workload": workload_name(variant, enc_ms, batch), "batch_per_gpu": batch, "n_samples": N_SAMPLES}


def pin_to_gpu_numa_node(local):
    """Binds this rank to the cores of its GPU's NUMA node (pinned host buffers allocated afterwards are then
    first-touched there): at N = 4 / 8 every rank used to sit on node 0 and the pinned H2D / D2H copies of the e2e
    path lost 5-6 %.  Best effort: any failure leaves the affinity alone."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if len(use) >= 4:
            os.sched_setaffinity(0, use)
            return {"numa_node": node, "cores": len(use), "how": "sysfs"}
    except Exception:
        pass
    try:
        # containers often hide the PCI devices' numa_node (-1): the driver's own table still has the affinity
        import re
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        txt = re.sub(r"\x1b\[[0-9;]*m", "", txt)
        lines = [ln for ln in txt.splitlines() if ln.strip()]
        head = [t.strip() for t in lines[0].split("\t")]
        col = head.index("CPU Affinity")
        row = next(ln for ln in lines[1:] if ln.split("\t")[0].strip() == f"GPU{local}")
        spec = [t.strip() for t in row.split("\t")][col]   # (the header's first cell is empty: columns line up)
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if len(use) >= 4 and use != allowed:
            os.sched_setaffinity(0, use)
            return {"cpu_affinity": spec, "cores": len(use), "how": "nvidia-smi topo"}
    except Exception:
        pass
    return None


def tf32_peak():
    """Dense TF32 tensor-core peak measured on this pool's B200 the way MEASURED_PEAKS.json measured bf16
    (torch.matmul 8192^3 with allow_tf32, best of 10 = burst, back to back for 4 s = sustained):
    profiles/r02_tf32_peak.json, written by scripts/measure_tf32_peak.py."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_tf32_peak.json")) as f:
            p = json.load(f)
        return float(p["tf32_tflops"]), float(p["tf32_tflops_sustained"]), "measured (profiles/r02_tf32_peak.json)"
    except Exception:
        return 1125.0, 1125.0, "fallback (half of the nominal dense bf16 2.25 PF)"


def gemm_flops(lengths, B, C=512, c=128):
    """2*M*K*N per forward (16 blocks) of the tcgen05 GEMM roles."""
    L0, Lb = lengths[0], lengths[-1]
    per_block = {"gemm_proj": 2 * C * c * L0, "gemm_res_conv": 2 * C * c * L0, "gemm_in_proj": 2 * 3 * C * C * Lb,
                 "gemm_out_proj": 2 * C * C * Lb, "gemm_fc1": 2 * 2 * C * C * Lb, "gemm_fc2": 2 * 2 * C * C * Lb}
    return {k: v * B * 16 for k, v in per_block.items()}


def eager_gpu_baseline(args, dev, B, steps=3):
    """"Eager PyTorch on B200" (SURVEY.md section 2a / BASELINE.md section 4): the reference's module graph as stock ATen /
    cuDNN / cuBLAS calls on this GPU - oracle/tdanet_oracle.py, the restatement pinned to the reference, moved to
    cuda - TF32 off and on, same weights, same batch.  A measured baseline like cpu_baseline, never the product."""
    from oracle import tdanet_oracle as O
    import tdanet_b200.look2hear.models as M
    kw = model_kwargs(args.enc_ms)
    torch.manual_seed(0)
    model = M.get(CLASSES[args.variant])(sample_rate=SR, **kw)
    sd = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(variant=args.variant, sample_rate=SR, **kw)
    x = (torch.randn(B, 1, N_SAMPLES, generator=torch.Generator().manual_seed(1234)) * 0.1).to(dev)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    out = {"unit": UNIT, "batch": B, "what": "oracle module graph as stock ATen/cuDNN/cuBLAS eager calls on this GPU"}
    try:
        for name, flag in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = flag
            with torch.no_grad():
                for _ in range(2):
                    O.forward(sd, x, cfg)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                ev0.record()
                for _ in range(steps):
                    O.forward(sd, x, cfg)
                ev1.record()
                torch.cuda.synchronize(dev)
            out[name] = B * (N_SAMPLES / SR) * steps / (ev0.elapsed_time(ev1) / 1e3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del sd, x
    torch.cuda.empty_cache()
    return out


def run_eager_gpu(args):
    """`--impl eager_gpu`: the eager-PyTorch-on-B200 arm as its own JSON line."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    e = eager_gpu_baseline(args, dev, args.batch, steps=max(1, args.steps))
    print(json.dumps({"impl": "eager_gpu", "metric": METRIC, "value": e["tf32"], "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": 2, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32 (TF32 matmul/cuDNN allowed)", "data": "synthetic",
                      "config": common_config(args.variant, args.enc_ms, args.batch), "eager_gpu": e}))


def write_detail(args, world, obj):
    """The full record (per-kernel table, every leg's config) goes to a file; stdout carries ONE compact line."""
    path = args.detail_out or os.path.join(ROOT, "gpurun_out", f"bench_detail_{args.variant}_{args.enc_ms}ms_{args.act_dtype}_{args.scaling}_n{world}.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(obj, f, indent=1)
        return os.path.relpath(path, ROOT)
    except OSError:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(variant, lengths, B, C=512, c=128):
    """Unique input + output bytes per forward for each kernel role (fp32, each tensor counted once per
    launch); the per-launch figures of DESIGN.md times the launches of one forward (16 blocks)."""
    L, depth, Lb, f = lengths, len(lengths), lengths[-1], 4
    blk = {}
    blk["gemm_proj"] = (L[0] * c + L[0] * C) * f
    blk["spp_dw0"] = 2 * L[0] * C * f
    blk["spp_dw_s2"] = sum((L[k - 1] + L[k]) * C for k in range(1, depth)) * f
    blk["pool_sum"] = (sum(L) + Lb) * C * f
    sl = sg = comb = first = 0
    for i in range(depth - 2, -1, -1):
        Lg = L[i - 1] if i == depth - 2 else L[i + 1]
        sl += L[i] * C * f
        sg += Lg * C * f
        if i == depth - 2:
            first = (L[i] + Lg + L[i]) * C * f
        else:
            comb += (L[i] + Lg + L[i]) * C * f
    blk["la_stats_local"], blk["la_stats_global"], blk["la_combine"], blk["la_combine_first"] = sl, sg, comb, first
    blk["gemm_res_conv"] = (L[0] * C + 3 * L[0] * c) * f
    blk["gemm_in_proj"] = Lb * 4 * C * f
    blk["gemm_out_proj"] = Lb * 2 * C * f
    blk["gemm_fc1"] = Lb * 3 * C * f
    blk["gemm_fc2"] = Lb * 3 * C * f
    blk["ffn_dw"] = Lb * 4 * C * f
    blk["attention"] = Lb * 4 * C * f
    return {k: v * B * 16 for k, v in blk.items()}


# ----------------------------------------------------------------------------- training step
TRAIN_BATCH = 8   # per GPU (BASELINE.json configs[3])


def train_algorithmic_bytes(lengths, B, C=512):
    """Unique input + output bytes per training step of the main backward roles (fp32, each tensor counted once
    per launch, 16 blocks); DESIGN.md section 8 lists what every launch reads and writes."""
    L, depth, Lb, f = lengths, len(lengths), lengths[-1], 4
    partner = (depth - 3 + depth) % depth
    la_a = la_dw = 0
    for i in range(depth - 1):
        Ll, Lg = L[i], (L[partner] if i == depth - 2 else L[i + 1])
        la_a += 4 * Ll + 11 * Lg       # G: 1R+2W (Lg); L: 2R+2W (Ll) + 1R+2W (Lg); F: 4R+1W (Lg)
        la_dw += 4 * Ll + 6 * Lg       # local: d_loc, raw_a, x_fused -> g_fused; global: 4 grads/raws + x_g -> g
    lgf_a = sum(4 * L[k] + 11 * Lb for k in range(depth - 1))
    lgf_dw = sum(4 * L[k] + 7 * Lb for k in range(depth - 1))
    blk = {
        "bwd_la_a": la_a, "bwd_la_dw": la_dw, "bwd_lgf_a": lgf_a, "bwd_lgf_dw": lgf_dw,
        "bwd_spp_dw_s2": sum(2 * L[k] + 3 * L[k - 1] for k in range(1, depth)),
        "bwd_spp_dw0": 4 * L[0],
        "bwd_pool": sum(2 * L[k] for k in range(depth)) + depth * Lb,
        "bwd_gln_stats": 2 * (sum(L) + L[0]),
    }
    return {k: v * C * f * B * 16 for k, v in blk.items()}


def train_targets(rank, B, seed=4321):
    tgt = torch.randn(B, 2, N_SAMPLES, generator=torch.Generator().manual_seed(seed + rank)) * 0.1
    return tgt.sum(1), tgt


def ncu_traffic(role):
    """DRAM bytes per launch of `role`'s kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json),
    or None when this round has no capture of it."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get(role)
        return (None, None) if e is None else (e["bytes_per_launch"], e["source"])
    except (OSError, ValueError, KeyError):
        return None, None


def run_longform_leg(args, dev, rank, world, barrier):
    """BASELINE.json configs[4]: 16 synthetic 60 s recordings, chunked like audio_test_css.py (segment 2 s, overlap
    0.25 -> 40 chunks each, the last one zero-padded), every chunk separated alone (attn_group 1), stitched on the
    device by cosine similarity.  The 16 recordings are split over the ranks (strong scaling, no collective).
    Returns the "longform" object of the JSON line (rank 0) or None."""
    import torch.distributed as dist
    import tdanet_b200.look2hear as look2hear
    n_streams, seconds = 16, 60.0
    if n_streams % world:
        return None
    mine = n_streams // world
    torch.manual_seed(0)
    model = getattr(look2hear.models, CLASSES[args.variant])(sample_rate=SR, **model_kwargs(args.enc_ms)).eval().to(dev)
    model.gemm_mode = args.gemm_mode
    model.act_dtype = args.act_dtype
    n = int(seconds * SR)
    wav_h = (torch.randn(mine, n, generator=torch.Generator().manual_seed(77 + rank)) * 0.1).pin_memory()
    wav = wav_h.to(dev)
    chunks = 160

    def once(src):
        return look2hear.system.separate_long(model, src, segment=2.0, overlap=0.25, max_chunks_per_call=chunks)

    for _ in range(3):
        out, swap = once(wav)
    K = max(3, min(args.steps, 10))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        out, swap = once(wav)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    out_h = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        o, _ = once(wav_h.to(dev, non_blocking=True))
        out_h.copy_(o, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    del model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    audio = n_streams * seconds * K
    return {"metric": "separated_audio_seconds_per_second", "value": audio / (ms / 1e3), "unit": UNIT,
            "ms_per_pass": ms / K, "scaling": "strong", "n_gpus": world,
            "config": {"workload": f"{CLASSES[args.variant]} {args.enc_ms} ms, long-form separation of 16 x 60 s recordings "
                                   "(BASELINE.json configs[4]): 40 chunks of 2 s per recording (overlap 0.25, last chunk padded), "
                                   "attention group 1, cosine-similarity stitch on the device",
                       "recordings_per_gpu": mine, "chunks_per_forward": chunks, "gemm_mode": args.gemm_mode,
                       "act_dtype": args.act_dtype},
            "e2e": {"value": audio / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_pass": mine * n * 4,
                    "d2h_bytes_per_pass": int(out.numel()) * 4},
            "stitched_samples": int(out.shape[-1]), "swapped_chunks": int(swap.sum().item())}


def run_two_ms_leg(args, dev, rank, world, barrier):
    """BASELINE.json configs[2]: the 2 ms encoder (latent 4010 .. 251), 64 mixtures split over the GPUs (strong
    scaling, no collective; every shard attends within itself like a per-shard reference run)."""
    import torch.distributed as dist
    import tdanet_b200.look2hear.models as M
    B = 64 // world
    torch.manual_seed(0)
    model = M.get(CLASSES[args.variant])(sample_rate=SR, **model_kwargs(2)).eval().to(dev)
    model.gemm_mode, model.act_dtype = args.gemm_mode, args.act_dtype
    eng, weights = model.engine, model._weights()
    x = (torch.randn(64, 1, N_SAMPLES, generator=torch.Generator().manual_seed(1234)) * 0.1)[rank * B:(rank + 1) * B]
    x = x.squeeze(1).contiguous().to(dev)
    K = max(5, min(args.steps, 10))
    with torch.no_grad():
        for _ in range(3):
            eng.forward_graphed(weights, x)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(K):
            eng.forward_graphed(weights, x)
        ev1.record()
        barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    del model, x
    eng._ws.clear()
    eng._graphs.clear()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, _ = measured_peaks()
    model_gb = {"best": 3.859, "fork": 3.892}[args.variant] / (2 if args.act_dtype == "bf16" else 1)
    value = 64 * 2.0 * K / (ms / 1e3)
    return {"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms / K, "steps": K, "scaling": "strong",
            "n_gpus": world, "batch_per_gpu": B, "global_batch": 64,
            "config": {"workload": workload_name(args.variant, 2, B)},
            "frac_of_survey_roofline": round(value / world / (peak / model_gb * 2.0), 4)}


def run_train_leg(args, dev, rank, world, local, barrier):
    """BASELINE.json configs[3]: full training step (forward, PIT SI-SDR loss, backward, gradient all-reduce,
    clip 5.0, Adam) of the 4 ms / 16-block TDANetBest at batch 8 per GPU.  Returns the "train" object of the
    JSON line (rank 0) or None."""
    import torch.distributed as dist
    import tdanet_b200.look2hear as look2hear
    from tdanet_b200 import _lib
    B, W, K = TRAIN_BATCH, max(3, args.warmup), args.steps
    torch.manual_seed(0)
    model = getattr(look2hear.models, CLASSES[args.variant])(sample_rate=SR, **model_kwargs(args.enc_ms)).to(dev).train()
    model.gemm_mode = args.gemm_mode
    model.act_dtype = args.act_dtype     # bf16: the large activations kept for the backward pass are stored as bf16
    L = look2hear.losses
    ts = look2hear.system.TrainingStep(model, L.PITLossWrapper(L.pairwise_neg_sisdr, threshold_byloss=True),
                                       lr=1e-3, clip_grad_norm=5.0)
    mix_h, tgt_h = train_targets(rank, B)
    mix_h, tgt_h = mix_h.pin_memory(), tgt_h.pin_memory()
    mix, tgt = mix_h.to(dev), tgt_h.to(dev)
    step = ts.step if args.no_graph else ts.step_captured
    n0 = _lib.launch_count()
    ts.step(mix, tgt)
    launches = _lib.launch_count() - n0
    for _ in range(W):
        step(mix, tgt)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(K):
            loss = step(mix, tgt)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    # end to end: this step's mixtures and targets come from pinned host memory, the loss goes back to the host
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()
    min_d, tgt_d = torch.empty_like(mix), torch.empty_like(tgt)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        min_d.copy_(mix_h, non_blocking=True)
        tgt_d.copy_(tgt_h, non_blocking=True)
        loss_h.copy_(step(min_d, tgt_d).reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    final_loss = float(loss_h.item())
    prof = None
    if rank == 0:
        _lib.profile_enable(True)
        ts.params.zero_grad()
        ts.forward_backward(mix, tgt)
        prof = _lib.profile_dump()
        _lib.profile_enable(False)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        return None
    alg = train_algorithmic_bytes(model.engine.latent_lengths(N_SAMPLES)[0], B) if args.variant == "best" else {}
    peak, peak_src = measured_peaks()
    kernels = []
    for p in sorted(prof, key=lambda r: -r["ms"]):
        row = {"kernel": p["kernel"], "launches_per_step": p["launches"], "ms_per_step": round(p["ms"], 4)}
        if p["kernel"] in alg:
            row["alg_GB_per_step"] = round(alg[p["kernel"]] / 1e9, 3)
            row["achieved_GBps"] = round(alg[p["kernel"]] / 1e9 / (p["ms"] / 1e3), 1)
        kernels.append(row)
    top = next((r for r in kernels if "achieved_GBps" in r), None)
    roofline = None if top is None else {
        "bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_GBps"], "peak": peak, "unit": "GB/s",
        "frac": round(top["achieved_GBps"] / peak, 4), "traffic": ncu_traffic(top["kernel"])[0], "traffic_source": ncu_traffic(top["kernel"])[1], "peak_source": peak_src,
        "algorithmic_bytes_per_launch": int(alg[top["kernel"]] / max(1, top["launches_per_step"])),
        "avg_launch_ms": round(top["ms_per_step"] / max(1, top["launches_per_step"]), 5),
        "share_of_step": round(top["ms_per_step"] / max(1e-9, sum(k["ms_per_step"] for k in kernels)), 4),
        "how": "CUDA events around every launch of one un-graphed forward+backward after the timed region "
               "(launch gaps inflate the short kernels; the graph-replayed step is what `value` times)"}
    out = {
        "metric": "train_steps_per_second", "value": K / (ms / 1e3), "unit": "steps/s", "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "n_gpus": world,
        "config": {"workload": f"{CLASSES[args.variant]} {args.enc_ms} ms encoder, 16 blocks, full training step (forward, PIT SI-SDR, "
                               f"backward, gradient all-reduce, clip 5.0, Adam), batch {B} x 2 s per GPU "
                               "(BASELINE.json configs[3]), train mode: dropout 0.1 / DropPath 0.1 like the reference (Philox keep-masks drawn on the device every step)",
                   "dropout": model.dropout, "drop_path": model.drop_path, "batch_per_gpu": B, "global_batch": B * world, "gemm_mode": args.gemm_mode,
                   "act_dtype": args.act_dtype,
                   "cuda_graph": not args.no_graph},
        "samples_per_second": world * B * K / (ms / 1e3),
        "e2e": {"value": K / (ms_e2e / 1e3), "unit": "steps/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": B * 3 * N_SAMPLES * 4, "d2h_bytes_per_step": 4},
        "gpu_launches_per_step": launches, "loss_after": final_loss, "clocks": clk.summary(), "roofline": roofline,
        # whole-step roofline: SURVEY.md 8(d) - a training step moves 3x the forward's bytes per mixture
        "step_frac": round((K / (ms / 1e3)) / (peak / (3 * {("best", 4): 1.936, ("best", 2): 3.859, ("fork", 4): 1.952, ("fork", 2): 3.892}[(args.variant, args.enc_ms)] * B)), 4),
        "profiled_ms_per_step": round(sum(k["ms_per_step"] for k in kernels), 3),
        "kernels": kernels,
    }
    if world == 1 and not args.skip_cpu:
        out["cpu_baseline"] = cpu_train_baseline(args)
    return out


def cpu_train_baseline(args):
    """Oracle port + autograd + torch clip/Adam on this box's host cores: one step at batch 2, scaled to batch 8."""
    from oracle import tdanet_oracle as O
    import tdanet_b200.look2hear.models as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = model_kwargs(args.enc_ms)
    torch.manual_seed(0)
    model = M.get(CLASSES[args.variant])(sample_rate=SR, **kw)
    sd = {k: (torch.nn.Parameter(v.detach().clone()) if "pos_enc.pe" not in k else v) for k, v in model.state_dict().items()}
    params = [v for k, v in sd.items() if "pos_enc.pe" not in k]
    opt = torch.optim.Adam(params, lr=1e-3)
    cfg = O.OracleConfig(variant=args.variant, sample_rate=SR, **kw)
    bs = 2
    mix, tgt = train_targets(0, bs)
    t0 = time.perf_counter()
    opt.zero_grad()
    loss = O.pit_loss(O.forward(sd, mix.unsqueeze(1), cfg), tgt, "sisdr", True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
    opt.step()
    dt = time.perf_counter() - t0
    return {"value": 1.0 / (dt * TRAIN_BATCH / bs), "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"one full step at batch {bs} (fp32 eager + autograd, {cores} threads) took {dt:.2f} s; "
                      f"scaled by {TRAIN_BATCH // bs} to the batch-{TRAIN_BATCH} step"}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference algorithm (oracle port of the PyTorch modules, fp32 eager) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import tdanet_oracle as O
    import tdanet_b200.look2hear.models as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = model_kwargs(args.enc_ms)
    torch.manual_seed(0)
    model = M.get(CLASSES[args.variant])(sample_rate=SR, **kw)    # parameter container only (CPU)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(variant=args.variant, sample_rate=SR, **kw)
    x1 = torch.randn(1, 1, N_SAMPLES, generator=torch.Generator().manual_seed(1234)) * 0.1
    with torch.no_grad():
        t0 = time.perf_counter()
        O.forward(sd, x1, cfg)
        t1 = time.perf_counter() - t0
    # The step is the full 64-mixture batch of the workload (batch-axis attention makes a smaller batch a different
    # attention problem, and the CPU is ~3x slower per mixture at batch 64 than at batch 8).  What is bounded is the
    # number of steps: one warm-up step is timed, then as many of the requested steps as fit in ~4 minutes (at least 3).
    bs = int(max(1, min(args.ref_batch, args.batch)))
    x = torch.randn(bs, 1, N_SAMPLES, generator=torch.Generator().manual_seed(1234)) * 0.1
    with torch.no_grad():
        t0 = time.perf_counter()
        O.forward(sd, x, cfg)
        t_step = time.perf_counter() - t0
        warm = 1
        while warm < args.warmup and (warm + 1 + args.steps) * t_step < 120.0:
            O.forward(sd, x, cfg)
            warm += 1
        steps = int(max(min(3, args.steps), min(args.steps, 240.0 // max(t_step, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(steps):
            O.forward(sd, x, cfg)
        dt = time.perf_counter() - t0
    requested_steps, args.steps = args.steps, steps
    value = bs * (N_SAMPLES / SR) * args.steps / dt
    sample = f"{bs} x 2 s mixtures per step (the workload's batch is {args.batch}), {args.steps} steps after {warm} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args.variant, args.enc_ms, args.batch),
        "impl_config": {"batch_per_step": bs, "steps_requested": requested_steps,
                        "note": "the reference is pure PyTorch and cannot travel to the GPU box (sources may not be "
                                "copied); this arm times oracle/tdanet_oracle.py, the CPU restatement pinned to the "
                                "reference by tests/golden, on all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- this framework
def run_ours(args):
    import torch.distributed as dist
    import tdanet_b200.look2hear.models as M
    from tdanet_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().tdanet_device_supported(local))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    if args.train_only:
        train = run_train_leg(args, dev, rank, world, local, barrier)
        if rank == 0:
            print(json.dumps(train))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.scaling == "strong" and args.batch % world:
        raise SystemExit(f"--scaling strong: batch {args.batch} does not split over {world} GPUs")
    B, W, K = (args.batch // world if args.scaling == "strong" else args.batch), max(3, args.warmup), args.steps
    kw = model_kwargs(args.enc_ms)
    torch.manual_seed(0)
    model = M.get(CLASSES[args.variant])(sample_rate=SR, **kw).eval().to(dev)
    model.gemm_mode = args.gemm_mode
    model.act_dtype = args.act_dtype
    eng, weights = model.engine, model._weights()
    # each rank separates its own shard of the global batch (weak: B mixtures per GPU; strong: args.batch / world)
    x_host = (torch.randn(B, 1, N_SAMPLES, generator=torch.Generator().manual_seed(1234 + rank)) * 0.1).pin_memory()
    x_dev = x_host.to(dev).squeeze(1).contiguous()
    lengths, _, _ = eng.latent_lengths(N_SAMPLES)

    def step():
        if args.no_graph:
            return eng.forward(weights, x_dev)
        return eng.forward_graphed(weights, x_dev)

    with torch.no_grad():
        n0 = _lib.launch_count()
        eng.forward(weights, x_dev)
        launches_per_step = _lib.launch_count() - n0
        for _ in range(W):
            step()
        # ---- timed region: K steps, inputs resident in HBM
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            barrier()
            ev0.record()
            for _ in range(K):
                step()
            ev1.record()
            barrier()
        ms = ev0.elapsed_time(ev1)
        # ---- end to end through the public API with host buffers
        out_host = torch.empty(B, 2, N_SAMPLES, dtype=torch.float32).pin_memory()
        model.use_cuda_graph = not args.no_graph
        xin = torch.empty(B, 1, N_SAMPLES, device=dev)
        for _ in range(2):
            xin.copy_(x_host, non_blocking=True)
            out_host.copy_(model(xin), non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            xin.copy_(x_host, non_blocking=True)        # H2D of this step's mixtures (pinned)
            est = model(xin)                            # the call a user makes
            out_host.copy_(est, non_blocking=True)      # D2H of the separated sources
            torch.cuda.current_stream().synchronize()   # the result is consumed on the host every step
        e1.record()
        barrier()
        ms_e2e_seq = e0.elapsed_time(e1)
        # the same K steps through look2hear.system.separate_pipelined: every step still copies its mixtures in from
        # pinned host memory and its sources back out, but the copies of neighbouring steps overlap the forward
        import tdanet_b200.look2hear.system as L2S
        outs2 = [out_host, torch.empty_like(out_host).pin_memory()]
        L2S.separate_pipelined(model, [x_host] * 3, outs2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L2S.separate_pipelined(model, [x_host] * K, outs2)   # returns after the last device-to-host copy
        e1.record()
        e1.synchronize()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        model.use_cuda_graph = False
        # ---- per-kernel durations (CUDA events around every launch, 2 un-graphed steps)
        prof = None
        if rank == 0:
            _lib.profile_enable(True)
            for _ in range(2):
                eng.forward(weights, x_dev)
            prof = _lib.profile_dump()
            _lib.profile_enable(False)

    t = torch.tensor([ms, ms_e2e, ms_e2e_seq], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks
    ms, ms_e2e, ms_e2e_seq = t.tolist()
    train = None
    if not args.skip_train:
        del x_dev, xin
        model._engine._ws.clear()
        model._engine._graphs.clear()
        torch.cuda.empty_cache()
        train = run_train_leg(args, dev, rank, world, local, barrier)
    longform = None
    if not args.skip_longform:
        model._engine._ws.clear()
        model._engine._graphs.clear()
        torch.cuda.empty_cache()
        longform = run_longform_leg(args, dev, rank, world, barrier)
    two_ms = None
    if not args.skip_2ms and args.enc_ms == 4 and 64 % world == 0:
        model._engine._ws.clear()
        model._engine._graphs.clear()
        torch.cuda.empty_cache()
        two_ms = run_two_ms_leg(args, dev, rank, world, barrier)
    eager = None
    if rank == 0 and world == 1 and not args.skip_eager:
        torch.cuda.empty_cache()
        eager = eager_gpu_baseline(args, dev, B)
    audio_s = world * B * (N_SAMPLES / SR) * K
    value = audio_s / (ms / 1e3)
    e2e = audio_s / (ms_e2e / 1e3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg = algorithmic_bytes(args.variant, lengths, B)
        if args.act_dtype == "bf16":
            # bf16 storage halves every large activation (the statistics, the 128-channel residual stream and the
            # bottom-scale tensors stay fp32): the streaming roles move half the bytes of the fp32 model
            half = ("gemm_proj", "spp_dw0", "spp_dw_s2", "la_stats_local", "la_stats_global", "la_combine", "la_combine_first")
            alg = {k: (v // 2 if k in half else v) for k, v in alg.items() if k in half}
        flops = gemm_flops(lengths, B)
        tf_burst, tf_sust, tf_src = tf32_peak()
        kernels = []
        for p in sorted(prof, key=lambda r: -r["ms"]):
            per_step_ms = p["ms"] / 2
            row = {"kernel": p["kernel"], "launches_per_step": p["launches"] // 2, "ms_per_step": round(per_step_ms, 4)}
            if p["kernel"] in alg:
                row["alg_GB_per_step"] = round(alg[p["kernel"]] / 1e9, 4)
                row["achieved_GBps"] = round(alg[p["kernel"]] / 1e9 / (per_step_ms / 1e3), 1)
                row["hbm_frac"] = round(row["achieved_GBps"] / peak, 4)
            if p["kernel"] in flops:      # the contraction kernels are reported on the tensor axis too
                row["achieved_TFLOPs"] = round(flops[p["kernel"]] / 1e12 / (per_step_ms / 1e3), 1)
                row["tensor_frac_of_tf32_sustained"] = round(row["achieved_TFLOPs"] / tf_sust, 4)
            kernels.append(row)
        top = next((r for r in kernels if "achieved_GBps" in r), None)
        prof_total = sum(r["ms_per_step"] for r in kernels)
        roofline = None if top is None else {
            "bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_GBps"], "peak": peak, "unit": "GB/s",
            "frac": round(top["achieved_GBps"] / peak, 4), "traffic": ncu_traffic(top["kernel"])[0], "traffic_source": ncu_traffic(top["kernel"])[1], "peak_source": peak_src,
            "algorithmic_bytes_per_launch": int(alg[top["kernel"]] / max(1, top["launches_per_step"])),
            "avg_launch_ms": round(top["ms_per_step"] / max(1, top["launches_per_step"]), 5),
            "share_of_step": round(top["ms_per_step"] / prof_total, 4),
            "how": "CUDA events around every launch on the launch stream, 2 un-graphed steps after the timed region",
        }
        # whole-step roofline against SURVEY.md §8(d)'s byte model (1.936 GB / mixture at 4 ms, fp32)
        model_gb = {("best", 4): 1.936, ("best", 2): 3.859, ("fork", 4): 1.952, ("fork", 2): 3.892}[(args.variant, args.enc_ms)]
        if args.act_dtype == "bf16":
            model_gb /= 2
        step_roofline = {"survey_bytes_per_mixture_GB": model_gb, "hbm_roofline_audio_s_per_s": round(peak / model_gb * 2.0, 1),
                         "frac_of_survey_roofline": round(value / world / (peak / model_gb * 2.0), 4)}
        cpu = cpu_baseline(args) if world == 1 and not args.skip_cpu else None
        dtype = ("bf16 storage, f32 arithmetic" if args.act_dtype == "bf16" else
                 "f32 (tf32 GEMMs)" if args.gemm_mode != "fp32" else "f32")
        detail = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": common_config(args.variant, args.enc_ms, B),
            "impl_config": {"gemm_mode": args.gemm_mode, "act_dtype": args.act_dtype, "cuda_graph": not args.no_graph,
                            "deterministic_statistics": bool(getattr(args, "deterministic", False)),
                            "global_batch": B * world, "numa": numa,
                            "l2": "no flush needed: one step streams a 1.7 GB workspace, 13x the 126 MB L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * N_SAMPLES * 4, "d2h_bytes_per_step": B * 2 * N_SAMPLES * 4,
                    "ms_per_step": ms_e2e / K,
                    "api": "look2hear.system.separate_pipelined(model, pinned host batches): every step copies its mixtures "
                           "in and its sources out; the copies of neighbouring steps overlap the forward",
                    "sequential": {"value": audio_s / (ms_e2e_seq / 1e3), "ms_per_step": ms_e2e_seq / K,
                                   "api": "xin.copy_(pinned); model(xin); out_pinned.copy_(est); synchronize - per step"}},
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clk.summary(),
            "roofline": roofline,
            "step_roofline": step_roofline,
            "tf32_peak": {"burst_TFLOPs": tf_burst, "sustained_TFLOPs": tf_sust, "source": tf_src},
            "kernels": kernels,
        }
        if cpu is not None:
            detail["cpu_baseline"] = cpu
        if eager is not None:
            detail["eager_gpu"] = eager
        if train is not None:
            detail["train"] = train
        if longform is not None:
            detail["longform"] = longform
        if two_ms is not None:
            detail["two_ms"] = two_ms
        detail_path = write_detail(args, world, detail)
        # ---- the ONE line on stdout: every contract key, compact (the driver keeps ~1.5 kB tails)
        r4 = lambda v: None if v is None else round(float(v), 4)   # noqa: E731
        r1 = lambda v: None if v is None else round(float(v), 1)   # noqa: E731
        out = {
            "metric": METRIC, "value": r1(value), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": r4(ms / K), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": detail["config"],
            "e2e": {"value": r1(e2e), "unit": UNIT, "h2d_bytes_per_step": B * N_SAMPLES * 4,
                    "d2h_bytes_per_step": B * 2 * N_SAMPLES * 4},
            "gpu_launches": launches_per_step * K,
            "clocks": {k: clk.summary()[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
            "roofline": None if roofline is None else {
                "bound": "hbm", "kernel": roofline["kernel"], "achieved": roofline["achieved"], "peak": peak, "unit": "GB/s",
                "frac": roofline["frac"], "traffic": roofline["traffic"] if args.act_dtype == "fp32" else None,
                "share": roofline["share_of_step"]},
            "step_frac": step_roofline["frac_of_survey_roofline"],
        }
        if cpu is not None:
            out["cpu_baseline"] = {"value": r4(cpu["value"]), "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
                                   "sample": cpu["sample_short"]}
        if eager is not None:
            out["eager_gpu"] = {"fp32": r1(eager["fp32"]), "tf32": r1(eager["tf32"])}
        if train is not None:
            out["train"] = {"value": round(train["value"], 2), "unit": "steps/s", "ms": round(train["ms_per_step"], 3),
                            "e2e": round(train["e2e"]["value"], 2), "batch_per_gpu": TRAIN_BATCH,
                            "frac": None if not train.get("roofline") else train["roofline"]["frac"],
                            "step_frac": train.get("step_frac")}
        if longform is not None:
            out["longform"] = {"value": r1(longform["value"]), "e2e": r1(longform["e2e"]["value"]), "scaling": "strong"}
        if two_ms is not None:
            out["two_ms"] = {"value": r1(two_ms["value"]), "scaling": "strong", "batch_per_gpu": two_ms["batch_per_gpu"]}
        if detail_path:
            out["detail"] = detail_path
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """Oracle port of the reference on this box's host cores, bounded sample (about 10-30 s of CPU work)."""
    from oracle import tdanet_oracle as O
    import tdanet_b200.look2hear.models as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = model_kwargs(args.enc_ms)
    torch.manual_seed(0)
    model = M.get(CLASSES[args.variant])(sample_rate=SR, **kw)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(variant=args.variant, sample_rate=SR, **kw)
    x = torch.randn(8, 1, N_SAMPLES, generator=torch.Generator().manual_seed(1234)) * 0.1
    with torch.no_grad():
        t0 = time.perf_counter()
        O.forward(sd, x[:1], cfg)                         # warm-up, and sizes the sample
        t1 = time.perf_counter() - t0
        bs = int(max(1, min(8, 12.0 // max(t1, 1e-3))))
        t0 = time.perf_counter()
        O.forward(sd, x[:bs], cfg)
        dt = time.perf_counter() - t0
    return {"value": bs * 2.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample_short": f"1 forward of {bs} of the {args.batch} mixtures, fp32 eager",
            "sample": f"one forward of {bs} x 2 s mixtures (of the {args.batch}-mixture step) after one 1-mixture warm-up, fp32 eager, {cores} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager_gpu"])
    ap.add_argument("--variant", default="best", choices=list(CLASSES))
    ap.add_argument("--enc-ms", type=int, default=4, choices=[2, 4])
    ap.add_argument("--batch", type=int, default=64, help="mixtures per GPU per step")
    ap.add_argument("--gemm-mode", default="tf32", choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--act-dtype", default="fp32", choices=["fp32", "bf16"],
                    help="storage of the large activations (bf16 = the bf16-mode tolerance of BASELINE.json)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--train-only", action="store_true", help="print the training-step leg as the JSON line (profiling aid)")
    ap.add_argument("--skip-train", action="store_true", help="skip the training-step leg (BASELINE.json configs[3])")
    ap.add_argument("--skip-longform", action="store_true", help="skip the long-form leg (BASELINE.json configs[4])")
    ap.add_argument("--ref-batch", type=int, default=64, help="upper bound of mixtures per step for --impl reference")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch mixtures per GPU; strong: --batch mixtures split over the GPUs (BASELINE.json configs[2])")
    ap.add_argument("--skip-eager", action="store_true", help="skip the eager-PyTorch-on-B200 baseline (N=1 only)")
    ap.add_argument("--skip-2ms", action="store_true", help="skip the strong-scaled 2 ms leg (BASELINE.json configs[2])")
    ap.add_argument("--detail-out", default="", help="where the full record goes (default gpurun_out/bench_detail_*.json)")
    ap.add_argument("--deterministic", action="store_true",
                    help="our arm with the deterministic-statistics mode on (bit-reproducible inference; measures its cost)")
    args = ap.parse_args()
    if args.deterministic and args.impl == "ours":
        from tdanet_b200 import _lib
        _lib.set_deterministic(True)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager_gpu":
        run_eager_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
