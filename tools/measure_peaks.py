#!/usr/bin/env python3
"""Measured roofline denominators the driver's MEASURED_PEAKS.json does not carry.

MEASUREMENT ONLY -- never on the product path.  Times cuBLAS through torch.matmul, with the same
protocol the driver used for its bf16 figure (MEASURED_PEAKS.json "how"): 8192^3, 2*N^3 FLOP,
best of 10 launches (burst) and back-to-back launches for 4 s (sustained), CUDA events.

    python tools/measure_peaks.py [--json profiles/r02_measured_peaks.json]

bench.py calls measure_tf32() in-process (before any timed region) when no committed figure is
found, and names the source in roofline.peak_source.
"""
import argparse
import json
import time


def _time_matmul(dtype, n=8192, sustained_s=4.0):
    import torch
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    flops = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = flops / (best * 1e-3) / 1e12
    # sustained: back to back for ~sustained_s seconds
    reps = max(10, int(sustained_s / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    return burst, sustained


def measure_tf32(sustained_s=4.0):
    """-> {"tf32_tflops": burst, "tf32_tflops_sustained": ...} from cuBLAS TF32 (fp32 tensors,
    allow_tf32) at 8192^3."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        burst, sustained = _time_matmul(torch.float32, sustained_s=sustained_s)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return {"tf32_tflops": burst, "tf32_tflops_sustained": sustained,
            "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32) 8192^3, 2*N^3: best of 10 "
                   "(burst) and back to back for %.0f s (sustained), CUDA events" % sustained_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--bf16", action="store_true", help="also re-measure bf16 with this script")
    args = ap.parse_args()
    import torch
    out = measure_tf32()
    if args.bf16:
        b, s = _time_matmul(torch.bfloat16)
        out.update({"bf16_tflops": b, "bf16_tflops_sustained": s})
        b, s = _time_matmul(torch.float16)
        out.update({"fp16_tflops": b, "fp16_tflops_sustained": s})
    out["gpu_name"] = torch.cuda.get_device_name(0)
    out["torch"] = torch.__version__
    out["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
    print(json.dumps(out))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
