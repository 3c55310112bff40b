"""world_size-2 checks of the multi-process host logic on CPU (gloo): rendezvous helpers used by
bench.py / replica.py (unique-id broadcast, max-over-ranks timing, batch sharding) and the
rank-0-only behaviour of the reference arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from simple_tensorflow_b200 import replica
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
payload = replica.broadcast_bytes(bytes(range(128)) if rank == 0 else None, src=0)
assert payload == bytes(range(128)), payload
assert replica.max_over_ranks(1.0 + rank) == float(world)
lo, hi = replica.shard_batch(32768, world, rank)
assert hi - lo == 32768 // world and lo == rank * (32768 // world)
sizes = [replica.shard_batch(10, 3, r) for r in range(3)]
assert sizes == [(0, 4), (4, 7), (7, 10)], sizes
dist.barrier()
print("rank %%d ok" %% rank)
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
        capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "1"], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
