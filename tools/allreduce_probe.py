"""NCCL all-reduce latency probe through the C ABI (b200_nccl_all_reduce), one rank per GPU.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_probe.py

Prints per message size the device time of one in-place fp32 ncclAvg all-reduce (CUDA events on
the launching stream, max over ranks) and the bus bandwidth 2(N-1)/N * bytes / t.  Not part of the
product or the test-suite; the numbers explain bench.py's N>1 step time.
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
    rows = []
    for nbytes in (4 << 10, 256 << 10, 1 << 20, 4 << 20, 12599296, 64 << 20):
        n = nbytes // 4
        buf = torch.ones(n, device="cuda", dtype=torch.float32)
        iters = 50

        def once():
            _lib.check(L.b200_nccl_all_reduce(_lib.DT_FLOAT, buf.data_ptr(), buf.data_ptr(), n, 1,
                                              comm, stream))
        for _ in range(5):
            once()
        _lib.check(L.b200_stream_synchronize(stream))
        dist.barrier()
        _lib.check(L.b200_event_record(e0, stream))
        for _ in range(iters):
            once()
        _lib.check(L.b200_event_record(e1, stream))
        _lib.check(L.b200_stream_synchronize(stream))
        ms = ctypes.c_float()
        _lib.check(L.b200_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        us = replica.max_over_ranks(ms.value * 1e3 / iters)
        assert abs(float(buf[0].item()) - 1.0) < 1e-6
        rows.append({"bytes": nbytes, "us": us,
                     "busbw_gbs": 2.0 * (world - 1) / world * nbytes / (us * 1e-6) / 1e9})
        if rank == 0:
            print("%10d B  %8.1f us  busbw %7.1f GB/s" % (nbytes, us, rows[-1]["busbw_gbs"]), flush=True)
    # ---- the NVLink peer-memory kernel (b200_peer_all_reduce), several grid sizes
    arena = ctypes.c_void_p()
    rc = L.b200_peer_arena_create(comm, rank, world, 64 << 20, ctypes.byref(arena))
    if rc != 0:
        if rank == 0:
            print("peer arena unavailable:", L.b200_last_error().decode(), flush=True)
    else:
        data = L.b200_peer_arena_data(arena)
        for ctas in (32, 64, 128, 256):
            for nbytes in (4 << 10, 4 << 20, 12599296):
                n = nbytes // 4
                ones = torch.ones(n, device="cuda", dtype=torch.float32)
                _lib.check(L.b200_memcpy_d2d_async(data, ones.data_ptr(), nbytes, stream))

                def once_peer():
                    _lib.check(L.b200_peer_all_reduce(arena, _lib.DT_FLOAT, 0, n, 1, ctas, stream))
                for _ in range(5):
                    once_peer()
                _lib.check(L.b200_stream_synchronize(stream))
                dist.barrier()
                _lib.check(L.b200_event_record(e0, stream))
                for _ in range(50):
                    once_peer()
                _lib.check(L.b200_event_record(e1, stream))
                _lib.check(L.b200_stream_synchronize(stream))
                ms = ctypes.c_float()
                _lib.check(L.b200_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
                us = replica.max_over_ranks(ms.value * 1e3 / 50)
                back = torch.empty(4, device="cuda", dtype=torch.float32)
                _lib.check(L.b200_memcpy_d2d_async(back.data_ptr(), data, 16, stream))
                _lib.check(L.b200_stream_synchronize(stream))
                assert abs(float(back[0].item()) - 1.0) < 1e-6, back
                rows.append({"bytes": nbytes, "us": us, "peer_kernel_ctas": ctas or "all",
                             "busbw_gbs": 2.0 * (world - 1) / world * nbytes / (us * 1e-6) / 1e9})
                if rank == 0:
                    print("%10d B  %8.1f us  busbw %7.1f GB/s  (peer kernel, ctas=%s)"
                          % (nbytes, us, rows[-1]["busbw_gbs"], ctas or "all"), flush=True)
        dist.barrier()
        _lib.check(L.b200_peer_arena_destroy(arena))
    # ---- same sizes on NCCL symmetric windows (ncclMemAlloc + ncclCommWindowRegister, 2.27+)
    try:
        nccl = ctypes.CDLL("libnccl.so.2")
        nccl.ncclMemAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
        nccl.ncclCommWindowRegister.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
        nccl.ncclAllReduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        for nbytes in (4 << 10, 1 << 20, 4 << 20, 12599296):
            n = nbytes // 4
            size = (nbytes + 4095) // 4096 * 4096
            ptr, win = ctypes.c_void_p(), ctypes.c_void_p()
            rc = nccl.ncclMemAlloc(ctypes.byref(ptr), size)
            assert rc == 0, "ncclMemAlloc rc=%d" % rc
            rc = nccl.ncclCommWindowRegister(comm, ptr, size, ctypes.byref(win), 1)
            assert rc == 0, "ncclCommWindowRegister rc=%d" % rc
            _lib.check(L.b200_memset_async(ptr, 0, size, stream))

            def once_sym():
                rc = nccl.ncclAllReduce(ptr, ptr, n, 7, 4, comm, stream)
                assert rc == 0, rc
            for _ in range(5):
                once_sym()
            _lib.check(L.b200_stream_synchronize(stream))
            dist.barrier()
            _lib.check(L.b200_event_record(e0, stream))
            for _ in range(50):
                once_sym()
            _lib.check(L.b200_event_record(e1, stream))
            _lib.check(L.b200_stream_synchronize(stream))
            ms = ctypes.c_float()
            _lib.check(L.b200_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
            us = replica.max_over_ranks(ms.value * 1e3 / 50)
            rows.append({"bytes": nbytes, "us": us, "symmetric_window": True})
            if rank == 0:
                print("%10d B  %8.1f us  (symmetric window)" % (nbytes, us), flush=True)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print("symmetric-window probe failed:", repr(e), flush=True)
    if rank == 0 and len(sys.argv) > 1:
        json.dump({"world": world, "rows": rows}, open(sys.argv[1], "w"), indent=1)
    dist.barrier()
    _lib.check(L.b200_nccl_comm_destroy(comm))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
