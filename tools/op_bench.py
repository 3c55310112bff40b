"""Per-op roofline probe: every hot-path op of the C ABI at BASELINE config sizes.

  python tools/op_bench.py            -> table: device time (CUDA events), achieved GB/s or TFLOP/s,
                                         fraction of the measured peak (MEASURED_PEAKS.json)
  ncu ... python tools/op_bench.py --once   -> one launch per op for an ncu capture

Inputs rotate through enough independent buffer sets that each launch reads data evicted from
the 126 MB L2 (working set per op x sets > 256 MB).  Not part of the test-suite or the product.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device memory + events only)

from simple_tensorflow_b200 import _lib  # noqa: E402

L = _lib.load()
F32, BF16 = _lib.DT_FLOAT, _lib.DT_BFLOAT16
dev = torch.device("cuda:0")


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p["bf16_tflops"], "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


HBM, BF16_TF, PEAK_SRC = peaks()
try:  # measured cuBLAS TF32 burst (tools/measure_peaks.py on this pool's B200)
    TF32_TF = json.load(open(os.path.join(ROOT, "profiles", "r02_measured_peaks.json")))["tf32_tflops"]
except Exception:
    TF32_TF = 0.5 * BF16_TF


def t(*shape, dtype=torch.float32):
    return torch.empty(*shape, device=dev, dtype=dtype).uniform_(-1, 1) if dtype.is_floating_point \
        else torch.zeros(*shape, device=dev, dtype=dtype)


def time_op(fn, sets, iters):
    st = torch.cuda.current_stream().cuda_stream
    for i in range(3):
        fn(sets[i % len(sets)], st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(sets[i % len(sets)], st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="run only ops whose name contains this substring")
    args = ap.parse_args()
    iters = 1 if args.once else args.iters
    rows = []

    def add(name, bytes_or_flops, unit, mk, fn, nsets=None):
        if args.only and args.only not in name:
            return
        per = max(1, bytes_or_flops if unit == "GB/s" else 64 << 20)
        n = nsets or max(2, min(12, int((300 << 20) // min(per, 300 << 20)) + 1))
        sets = [mk() for _ in range(n)]
        us = time_op(fn, sets, iters) if not args.once else (fn(sets[0], torch.cuda.current_stream().cuda_stream), torch.cuda.synchronize(), 0.0)[2]
        if unit == "GB/s":
            ach = bytes_or_flops / (us * 1e-6) / 1e9 if us else 0
            peak = HBM
        else:
            ach = bytes_or_flops / (us * 1e-6) / 1e12 if us else 0
            peak = BF16_TF if "bf16" in name else TF32_TF
        rows.append(dict(op=name, us=us, achieved=ach, unit=unit, peak=peak, frac=ach / peak if peak else 0))
        if not args.once:
            print("%-46s %9.1f us %10.1f %-8s %5.1f%% of %s peak" % (name, us, ach, unit, 100 * ach / peak, PEAK_SRC), flush=True)
        del sets
        torch.cuda.empty_cache()

    R, C = 4096, 1024  # C2 activations
    n = R * C
    # ---- streaming ops (C2 shapes)
    add("BiasAdd f32 [4096,1024]", 2 * n * 4 + C * 4, "GB/s", lambda: (t(R, C), t(C), t(R, C)),
        lambda s, st: L.b200_bias_add(F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), R, C, st))
    add("Relu f32 [4096,1024]", 2 * n * 4, "GB/s", lambda: (t(R, C), t(R, C)),
        lambda s, st: L.b200_relu(F32, s[0].data_ptr(), s[1].data_ptr(), n, st))
    add("ReluGrad f32 [4096,1024]", 3 * n * 4, "GB/s", lambda: (t(R, C), t(R, C), t(R, C)),
        lambda s, st: L.b200_relu_grad(F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), n, st))
    wsb = L.b200_bias_add_grad_workspace_bytes(F32, R, C)
    add("BiasAddGrad f32 [4096,1024] (2 kernels)", n * 4 + C * 4, "GB/s",
        lambda: (t(R, C), t(C), torch.empty(wsb, device=dev, dtype=torch.uint8)),
        lambda s, st: L.b200_bias_add_grad(F32, s[0].data_ptr(), s[1].data_ptr(), R, C, s[2].data_ptr(), wsb, st))
    add("Softmax f32 [4096,1024]", 2 * n * 4, "GB/s", lambda: (t(R, C), t(R, C)),
        lambda s, st: L.b200_softmax(F32, s[0].data_ptr(), s[1].data_ptr(), R, C, 0, st))
    add("SoftmaxXent f32 [4096,1024]", 3 * n * 4 + R * 4, "GB/s", lambda: (t(R, C), t(R, C), t(R), t(R, C)),
        lambda s, st: L.b200_softmax_xent(F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), s[3].data_ptr(), R, C, st))
    add("Cast f32->bf16 [4096,1024] (truncate)", n * 6, "GB/s", lambda: (t(R, C), t(R, C, dtype=torch.bfloat16)),
        lambda s, st: L.b200_cast(F32, BF16, s[0].data_ptr(), s[1].data_ptr(), n, st))
    add("ArgMax f32 [4096,1024] axis 1", n * 4 + R * 8, "GB/s", lambda: (t(R, C), torch.zeros(R, device=dev, dtype=torch.int64)),
        lambda s, st: L.b200_argmax(F32, s[0].data_ptr(), s[1].data_ptr(), R, C, 1, st))
    add("ApplyGradientDescent f32 [1024,1024]", 3 * C * C * 4, "GB/s", lambda: (t(C, C), t(1), t(C, C)),
        lambda s, st: L.b200_apply_gradient_descent(F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), C * C, st))
    # ---- pooling (C3 pool1)
    N, H, W, CH = 512, 28, 28, 32
    ni, no = N * H * W * CH, N * 14 * 14 * CH
    add("MaxPool f32 [512,28,28,32] 2x2/2", (ni + no) * 4, "GB/s", lambda: (t(N, H, W, CH), t(N, 14, 14, CH)),
        lambda s, st: L.b200_max_pool(F32, s[0].data_ptr(), s[1].data_ptr(), N, H, W, CH, 14, 14, 2, 2, 2, 2, 0, 0, st))
    add("MaxPoolGrad f32 [512,28,28,32] 2x2/2", (2 * ni + no) * 4, "GB/s", lambda: (t(N, H, W, CH), t(N, 14, 14, CH), t(N, H, W, CH)),
        lambda s, st: L.b200_max_pool_grad(F32, s[0].data_ptr(), None, s[1].data_ptr(), s[2].data_ptr(), N, H, W, CH, 14, 14, 2, 2, 2, 2, 0, 0, st))
    # ---- GEMMs (C2 / C4 shapes)
    for name, dt, tdt in (("tf32", F32, torch.float32), ("bf16", BF16, torch.bfloat16)):
        for tag, m, nn, k, ta, tb in (("fwd  X.W", 4096, 1024, 1024, 0, 0), ("dX   dY.W^T", 4096, 1024, 1024, 0, 1),
                                      ("dW   X^T.dY", 1024, 1024, 4096, 1, 0)):
            ws = L.b200_matmul_workspace_bytes(dt, m, nn, k)
            add("MatMul %s %s %dx%dx%d" % (name, tag, m, nn, k), 2.0 * m * nn * k, "TFLOP/s",
                lambda: (t(k, m, dtype=tdt) if ta else t(m, k, dtype=tdt), t(nn, k, dtype=tdt) if tb else t(k, nn, dtype=tdt),
                         t(m, nn, dtype=tdt), torch.empty(max(ws, 1), device=dev, dtype=torch.uint8)),
                lambda s, st, m=m, nn=nn, k=k, ta=ta, tb=tb, dt=dt, ws=ws: L.b200_matmul(
                    dt, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), m, nn, k, ta, tb,
                    s[3].data_ptr() if ws else None, ws, st), nsets=6)
    # ---- BatchMatMul on the reference's own benchmark shapes (batch_matmul_op_test.cc:62-132:
    # BM_BatchMatmul(B, M, K, N, adj_x, adj_y)), fp32 graph on the TF32 tensor path
    for b, m, k, nn in ((1, 128, 1024, 1024), (8, 128, 1024, 1024), (32, 128, 1024, 1024),
                        (8, 256, 256, 256), (32, 256, 256, 256), (8, 1024, 1024, 1024),
                        (32, 1024, 1024, 1024), (8, 2048, 2048, 2048), (32, 10000, 200, 1),
                        (32, 1, 200, 10000)):
        add("BatchMatMul tf32 B%d %dx%dx%d (MxKxN)" % (b, m, k, nn), 2.0 * b * m * nn * k, "TFLOP/s",
            lambda b=b, m=m, k=k, nn=nn: (t(b, m, k), t(b, k, nn), t(b, m, nn)),
            lambda s, st, b=b, m=m, k=k, nn=nn: L.b200_batch_matmul(
                F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), b, m, nn, k, 0, 0, st), nsets=4)
    wsb16 = L.b200_bias_add_grad_workspace_bytes(BF16, R, C)
    add("BiasAddGrad bf16 [4096,1024]", n * 2 + C * 2, "GB/s",
        lambda: (t(R, C, dtype=torch.bfloat16), t(C, dtype=torch.bfloat16), torch.empty(wsb16, device=dev, dtype=torch.uint8)),
        lambda s, st: L.b200_bias_add_grad(BF16, s[0].data_ptr(), s[1].data_ptr(), R, C, s[2].data_ptr(), wsb16, st))
    add("SoftmaxXent bf16 [4096,1024]", 3 * n * 2 + R * 2, "GB/s",
        lambda: (t(R, C, dtype=torch.bfloat16), t(R, C, dtype=torch.bfloat16), t(R, dtype=torch.bfloat16), t(R, C, dtype=torch.bfloat16)),
        lambda s, st: L.b200_softmax_xent(BF16, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), s[3].data_ptr(), R, C, st))
    add("FusedMatMul tf32 fwd+bias+relu 4096x1024x1024", 2.0 * 4096 * 1024 * 1024, "TFLOP/s",
        lambda: (t(4096, 1024), t(1024, 1024), t(4096, 1024), t(1024)),
        lambda s, st: L.b200_fused_matmul(F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), 4096, 1024, 1024, 0, 0,
                                          s[3].data_ptr(), 1, None, st), nsets=6)
    add("FusedMatMul tf32 dX+relugrad 4096x1024x1024", 2.0 * 4096 * 1024 * 1024, "TFLOP/s",
        lambda: (t(4096, 1024), t(1024, 1024), t(4096, 1024), t(4096, 1024)),
        lambda s, st: L.b200_fused_matmul(F32, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), 4096, 1024, 1024, 0, 1,
                                          None, 0, s[3].data_ptr(), st), nsets=6)
    # ---- convolution (C3 conv2: 512x14x14x32 * 5x5x32x64 SAME)
    g = _lib.ConvGeometry(512, 14, 14, 32, 5, 5, 64, 14, 14, 1, 1, 2, 2)
    flops = 2.0 * 512 * 14 * 14 * 5 * 5 * 32 * 64
    for which, nm in ((0, "Conv2D"), (1, "Conv2DBackpropInput"), (2, "Conv2DBackpropFilter")):
        ws = L.b200_conv2d_workspace_bytes(F32, ctypes.byref(g), which)

        def mk(ws=ws):
            return (t(512, 14, 14, 32), t(5, 5, 32, 64), t(512, 14, 14, 64), torch.empty(max(ws, 1), device=dev, dtype=torch.uint8))

        def run(s, st, which=which, ws=ws):
            x, w, y, wk = s
            if which == 0:
                return L.b200_conv2d(F32, x.data_ptr(), w.data_ptr(), y.data_ptr(), ctypes.byref(g), wk.data_ptr(), ws, st)
            if which == 1:
                return L.b200_conv2d_backprop_input(F32, w.data_ptr(), y.data_ptr(), x.data_ptr(), ctypes.byref(g), wk.data_ptr(), ws, st)
            return L.b200_conv2d_backprop_filter(F32, x.data_ptr(), y.data_ptr(), w.data_ptr(), ctypes.byref(g), wk.data_ptr(), ws, st)
        add("%s tf32 LeNet conv2 batch 512" % nm, flops, "TFLOP/s", mk, run, nsets=2)
    if args.json and not args.once:
        json.dump({"peaks": {"hbm_gbs": HBM, "bf16_tflops": BF16_TF, "source": PEAK_SRC,
                             "tf32_tflops": TF32_TF,
                             "tf32_peak": "measured cuBLAS TF32 burst (profiles/r02_measured_peaks.json)"}, "rows": rows},
                  open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
