"""Dense TF32 tensor-core peak of this pool's B200, measured the way MEASURED_PEAKS.json measured bf16
(torch.matmul 8192^3, 2*N^3 flops; best of 10 = burst, back to back for 4 s = sustained), with
torch.backends.cuda.matmul.allow_tf32.  Writes profiles/r02_tf32_peak.json ("TF32 peak not
measured - measure it with the same method before quoting a fraction")."""
import json
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_tf32_peak.json")
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    flops = 2.0 * n ** 3
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) / 1e3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, reps = time.perf_counter(), 0
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = flops * reps / (e0.elapsed_time(e1) / 1e3) / 1e12
    # the same two numbers for bf16, as a cross-check against MEASURED_PEAKS.json
    ab, bb = a.bfloat16(), b.bfloat16()
    for _ in range(5):
        ab @ bb
    torch.cuda.synchronize()
    best16 = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ab @ bb
        e1.record()
        torch.cuda.synchronize()
        best16 = max(best16, flops / (e0.elapsed_time(e1) / 1e3) / 1e12)
    rec = {"tf32_tflops": round(best, 1), "tf32_tflops_sustained": round(sustained, 1), "bf16_tflops_check": round(best16, 1),
           "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": "torch.matmul fp32 8192^3 with allow_tf32 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained)"}
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
