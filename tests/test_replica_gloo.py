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


def test_reference_arm_line_follows_the_bench_contract():
    """`bench.py --impl reference` (the CPU oracle port on the host cores): ONE JSON line on
    stdout with the contract's keys, `impl: reference`, zero copy bytes, its own cpu_baseline."""
    env = dict(os.environ)
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "1", "--steps", "1", "--warmup", "1"], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "cpu_baseline"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 4096 * d["steps"] / (d["ms_per_step"] * d["steps"] / 1e3)) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
