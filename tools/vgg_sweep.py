"""BASELINE configs[4] (C5) roofline sweep: the 8 convolutions of a VGG-style stack (3x3, SAME,
stride 1, bf16 storage / fp32 accumulate), batch 32 per replica, input resolution in
{32, 64, 112, 224}; forward, input gradient and filter gradient of every layer through the C ABI.

    python tools/vgg_sweep.py [--json profiles/r02_vgg_sweep.json] [--sizes 32,64,112,224]

Device time with CUDA events (tools/op_bench.py protocol: rotating buffer sets larger than L2),
TFLOP/s = 2 * N*OH*OW*K*R*S*C / time, fraction of the measured bf16 cuBLAS burst peak.
Measurement tool; not part of the test-suite or the product.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import op_bench as ob  # noqa: E402  (helpers only: t(), time_op(), L, peaks)
from simple_tensorflow_b200 import _lib  # noqa: E402

L, BF16 = ob.L, ob.BF16
CHANNELS = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512), (512, 512)]
POOL_AFTER = {1, 3, 5, 7}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--sizes", default="32,64,112,224")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rows = []
    bt = torch.bfloat16
    for size in [int(s) for s in args.sizes.split(",")]:
        h = size
        for li, (c, k) in enumerate(CHANNELS):
            n = args.batch
            g = _lib.ConvGeometry(n, h, h, c, 3, 3, k, h, h, 1, 1, 1, 1)
            flops = 2.0 * n * h * h * k * 9 * c
            for which, nm in ((0, "fwd"), (1, "dX"), (2, "dW")):
                if which == 1 and li == 0:
                    continue  # no gradient into the image
                ws = L.b200_conv2d_workspace_bytes(BF16, ctypes.byref(g), which)
                per_set = (n * h * h * (c + k) * 2 + max(ws, 1))
                nsets = max(2, min(6, int((300 << 20) // max(per_set, 1)) + 1))

                def mk(ws=ws, n=n, h=h, c=c, k=k):
                    return (ob.t(n, h, h, c, dtype=bt), ob.t(3, 3, c, k, dtype=bt), ob.t(n, h, h, k, dtype=bt),
                            torch.empty(max(ws, 1), device=ob.dev, dtype=torch.uint8))

                def run(s, st, which=which, ws=ws, g=g):
                    x, w, y, wk = s
                    if which == 0:
                        return L.b200_conv2d(BF16, x.data_ptr(), w.data_ptr(), y.data_ptr(), ctypes.byref(g), wk.data_ptr(), ws, st)
                    if which == 1:
                        return L.b200_conv2d_backprop_input(BF16, w.data_ptr(), y.data_ptr(), x.data_ptr(), ctypes.byref(g), wk.data_ptr(), ws, st)
                    return L.b200_conv2d_backprop_filter(BF16, x.data_ptr(), y.data_ptr(), w.data_ptr(), ctypes.byref(g), wk.data_ptr(), ws, st)
                sets = [mk() for _ in range(nsets)]
                us = ob.time_op(run, sets, args.iters)
                del sets
                torch.cuda.empty_cache()
                tf = flops / (us * 1e-6) / 1e12
                row = dict(input=size, layer=li + 1, hw=h, cin=c, cout=k, op=nm, us=us, tflops=tf,
                           frac_of_bf16_burst=tf / ob.BF16_TF, gflop=flops / 1e9)
                rows.append(row)
                print("in %3d  conv%d %3dx%-3d %3d->%-3d %-3s %9.1f us %8.1f TFLOP/s  %5.1f%% of bf16 burst"
                      % (size, li + 1, h, h, c, k, nm, us, tf, 100 * tf / ob.BF16_TF), flush=True)
            if li in POOL_AFTER:
                h = max(1, h // 2)
    if args.json:
        json.dump({"peak_bf16_tflops": ob.BF16_TF, "peak_source": ob.PEAK_SRC, "batch": args.batch,
                   "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
