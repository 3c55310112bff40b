"""All-reduce latency of the gradient-arena kernels (NVLS multimem / peer-IPC two-shot) through the
C ABI, one rank per GPU:

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/arena_probe.py

Per message size and CTA cap: device time of one in-place fp32 average (CUDA events around 50
back-to-back calls on the launching stream, max over ranks).  B200TF_NVLS=0 selects the IPC kernel.
Measurement aid only (explains bench.py's N>1 step time); not part of the product or the tests.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from simple_tensorflow_b200 import _lib, replica  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.load()
    comm = replica.init_nccl_comm(L, rank, world, local)
    stream = ctypes.c_void_p()
    _lib.check(L.b200_stream_create(ctypes.byref(stream)))
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.b200_event_create(ctypes.byref(e0)))
    _lib.check(L.b200_event_create(ctypes.byref(e1)))
    arena = ctypes.c_void_p()
    rc = L.b200_peer_arena_create(comm, rank, world, 64 << 20, ctypes.byref(arena))
    if rc != 0:
        if rank == 0:
            print("arena unavailable:", L.b200_last_error().decode(), flush=True)
        return
    L.b200_peer_arena_backend.restype = ctypes.c_char_p
    backend = L.b200_peer_arena_backend(arena).decode()
    data = L.b200_peer_arena_data(arena)
    rows = []
    for ctas in (16, 32, 64, 128, 256):
        for nbytes in (4 << 10, 256 << 10, 1 << 20, 4 << 20, 12599296, 32 << 20):
            n = nbytes // 4
            ones = torch.ones(n, device="cuda", dtype=torch.float32)
            _lib.check(L.b200_memcpy_d2d_async(data, ones.data_ptr(), nbytes, stream))

            def once():
                _lib.check(L.b200_peer_all_reduce(arena, _lib.DT_FLOAT, 0, n, 1, ctas, stream))
            for _ in range(5):
                once()
            _lib.check(L.b200_stream_synchronize(stream))
            dist.barrier()
            _lib.check(L.b200_event_record(e0, stream))
            for _ in range(50):
                once()
            _lib.check(L.b200_event_record(e1, stream))
            _lib.check(L.b200_stream_synchronize(stream))
            ms = ctypes.c_float()
            _lib.check(L.b200_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
            us = replica.max_over_ranks(ms.value * 1e3 / 50)
            back = torch.empty(4, device="cuda", dtype=torch.float32)
            _lib.check(L.b200_memcpy_d2d_async(back.data_ptr(), data, 16, stream))
            _lib.check(L.b200_stream_synchronize(stream))
            assert abs(float(back[0].item()) - 1.0) < 1e-6, back
            rows.append({"backend": backend, "bytes": nbytes, "ctas": ctas, "us": us})
            if rank == 0:
                print("%-8s ctas %3d %10d B  %8.1f us  algbw %7.1f GB/s" %
                      (backend, ctas, nbytes, us, nbytes / (us * 1e-6) / 1e9), flush=True)
    dist.barrier()
    _lib.check(L.b200_peer_arena_destroy(arena))
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out", "r02"), exist_ok=True)
        json.dump({"world": world, "rows": rows},
                  open(os.path.join(ROOT, "gpurun_out", "r02", "arena_probe_%s_n%d.json" % (backend, world)), "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
