#!/usr/bin/env python
"""Measure the tensor-pipe peaks MEASURED_PEAKS.json does not carry, with the driver's own recipe (SURVEY.md s.8d,
BASELINE.md s.3: "a TF32 peak must be measured the same way if kind::tf32 MMA is used"): torch.matmul of 8192^3 (2 N^3 flops),
best of 10 (burst) and back to back for 4 s (sustained), for fp32 inputs with TF32 allowed and for fp16 inputs.

    python scripts/measure_peaks.py [--out gpurun_out/measured_tf32_peak.json]      (1 GPU; copy the result to profiles/)
"""
import argparse
import json
import time

import torch


def measure(dtype, tf32):
    n = 8192
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    flops = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
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
    return best, reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/measured_tf32_peak.json")
    a = ap.parse_args()
    tb, ts = measure(torch.float32, True)
    hb, hs = measure(torch.float16, False)
    out = {"tf32_tflops": tb, "tf32_tflops_sustained": ts, "fp16_tflops": hb, "fp16_tflops_sustained": hs,
           "gpu_name": torch.cuda.get_device_name(0), "torch": str(torch.__version__),
           "how": "torch.matmul 8192^3 (2 N^3): best of 10 (burst) and back to back for 4 s (sustained); fp32 inputs with "
                  "allow_tf32 = True (cuBLAS TF32 tensor cores) and fp16 inputs",
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out))
