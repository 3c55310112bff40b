"""Replica data-parallel plumbing: one process per GPU, one graph copy per process.

torch.distributed is only the rendezvous (it carries the 128-byte NCCL unique id and the
max-over-ranks of timings); the gradient exchange itself is the B200AllReduceN op calling
ncclAllReduce on the session's compute stream through the C ABI (b200_nccl_all_reduce_sum).
"""
import ctypes


def shard_batch(global_batch, world_size, rank):
    """Rows [lo, hi) of the global batch owned by `rank` (even split on dim 0; the remainder
    goes to the lowest ranks)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` to every rank of the default process group."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def max_over_ranks(value):
    """max of a python float over all ranks (the timing rule for multi-GPU numbers)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def init_nccl_comm(lib, rank, world_size, local_rank):
    """Create this rank's NCCL communicator through the C ABI; returns the ncclComm_t as int."""
    uid = ctypes.create_string_buffer(128)
    if rank == 0:
        rc = lib.b200_nccl_unique_id(uid)
        if rc != 0:
            raise RuntimeError(lib.b200_last_error().decode())
    payload = broadcast_bytes(uid.raw if rank == 0 else None, src=0)
    uid = ctypes.create_string_buffer(payload, 128)
    rc = lib.b200_set_device(local_rank)
    if rc != 0:
        raise RuntimeError(lib.b200_last_error().decode())
    comm = ctypes.c_void_p()
    rc = lib.b200_nccl_comm_init_rank(ctypes.byref(comm), world_size, uid, rank)
    if rc != 0:
        raise RuntimeError(lib.b200_last_error().decode())
    return comm.value
