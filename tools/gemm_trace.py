"""Developer probe: B200TF_GEMM_TRACE=1 python tools/gemm_trace.py -> the phase boundaries of CTA 0
of the tcgen05 GEMM (ns since kernel entry) for the MLP shapes, cold (first launch) and warm."""
import ctypes, os, sys
os.environ["B200TF_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_tensorflow_b200 import _lib

L = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for dt, tdt in ((_lib.DT_FLOAT, torch.float32), (_lib.DT_BFLOAT16, torch.bfloat16)):
    for m, n, k, ta, tb in ((4096, 1024, 1024, 0, 0), (4096, 1024, 1024, 0, 1), (1024, 1024, 4096, 1, 0)):
        a = torch.randn((k, m) if ta else (m, k), device=dev).to(tdt)
        b = torch.randn((n, k) if tb else (k, n), device=dev).to(tdt)
        c = torch.empty((m, n), device=dev, dtype=tdt)
        ws = L.b200_matmul_workspace_bytes(dt, m, n, k)
        w = torch.empty(max(ws, 1), device=dev, dtype=torch.uint8)
        for rep in range(3):
            _lib.check(L.b200_matmul(dt, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, ta, tb,
                                     w.data_ptr() if ws else None, ws, st))
        bias = torch.randn(n, device=dev).to(tdt)
        if not ta:
            for rep in range(2):
                _lib.check(L.b200_fused_matmul(dt, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, ta, tb,
                                               bias.data_ptr(), 1, None, st))
torch.cuda.synchronize()
