"""Developer probe: run the C-ABI MatMul on a B200 in every layout / dtype / tile variant and
print error vs an fp64 reference plus device time.  Not part of the product or the test-suite."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_tensorflow_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)

def run(dtype, m, n, k, ta, tb, batch=1, iters=0):
    tdt = torch.float32 if dtype == _lib.DT_FLOAT else torch.bfloat16
    a = torch.randn((batch, k, m) if ta else (batch, m, k), device=dev).to(tdt)
    b = torch.randn((batch, n, k) if tb else (batch, k, n), device=dev).to(tdt)
    c = torch.full((batch, m, n), float("nan"), device=dev, dtype=tdt)
    st = torch.cuda.current_stream().cuda_stream
    if batch == 1:
        nb = lib.b200_matmul_workspace_bytes(dtype, m, n, k)
        ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
        wsp = ws.data_ptr() if nb else None
        rc = lib.b200_matmul(dtype, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, int(ta), int(tb), wsp, nb, st)
    else:
        rc = lib.b200_batch_matmul(dtype, a.data_ptr(), b.data_ptr(), c.data_ptr(), batch, m, n, k, int(ta), int(tb), st)
    if rc:
        print("  rc", rc, lib.b200_last_error()); return
    torch.cuda.synchronize()
    A = a.double().transpose(1, 2) if ta else a.double()
    B = b.double().transpose(1, 2) if tb else b.double()
    ref = A @ B
    err = (c.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    nan = torch.isnan(c).sum().item()
    msg = f"dt={dtype} m={m} n={n} k={k} ta={int(ta)} tb={int(tb)} b={batch}: maxerr={err:.3e} rel={err/scale:.3e} nan={nan}"
    if iters:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            lib.b200_matmul(dtype, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, int(ta), int(tb), wsp, nb, st)
        e0.record()
        for _ in range(iters):
            lib.b200_matmul(dtype, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, int(ta), int(tb), wsp, nb, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        msg += f"  {ms*1e3:.1f} us  {2*m*n*k/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)

print(lib.b200_version(), "devices", lib.b200_device_count())
for dtype in (_lib.DT_FLOAT, _lib.DT_BFLOAT16):
    for ta in (False, True):
        for tb in (False, True):
            run(dtype, 256, 256, 256, ta, tb)
    run(dtype, 128, 128, 128, False, True)
    run(dtype, 200, 136, 72, False, False)      # ragged tiles, aligned strides
    run(dtype, 200, 136, 72, True, True)
    run(dtype, 512, 384, 200, False, False)
    run(dtype, 300, 130, 64, True, False)
    run(dtype, 4096, 1024, 1024, False, False, iters=20)
    run(dtype, 4096, 1024, 1024, False, True, iters=20)
    run(dtype, 1024, 1024, 4096, True, False, iters=20)
    run(dtype, 512, 512, 256, False, False, batch=4)
    run(dtype, 100, 60, 36, True, False, batch=3)
    run(dtype, 3, 5, 7, False, False)           # SIMT path
    run(dtype, 37, 53, 71, True, True)          # SIMT path (unaligned)
print("launches", lib.b200_launch_count())
