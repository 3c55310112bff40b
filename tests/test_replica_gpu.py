"""Replica data-parallel parity on real GPUs: world_size-2 training through
Session.run + the B200AllReduceN collective against the CPU oracle on the whole batch.
Needs >= 2 GPUs on the box (skipped otherwise); the rendezvous host logic alone is covered on CPU
by tests/test_replica_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gpu_count():
    from simple_tensorflow_b200 import _lib
    return _lib.load().b200_device_count()


@pytest.mark.parametrize("bucket,overlap,pack,peer", [
    ("default", "0", "1", "1"),   # one NVLink peer-memory all-reduce of the gradient arena
    ("default", "0", "1", "0"),   # ... the same through NCCL
    ("1048576", "1", "1", "1"),   # per-layer buckets on the collective stream, peer kernel
    ("1048576", "1", "1", "0"),   # ... NCCL
    ("1048576", "1", "0", "0"),   # grouped in-place collectives on the first (unlearned) step
    ("1048576", "0", "1", "1"),   # per-layer buckets on the compute stream
    ("2048", "1", "1", "1"),      # tiny buckets: every gradient its own collective
])
def test_two_replica_training_matches_full_batch_oracle(bucket, overlap, pack, peer):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, REPLICA_TEST_BUCKET=bucket, B200TF_COLLECTIVE_OVERLAP=overlap,
               B200TF_ALLREDUCE_PACK=pack, B200TF_PEER_ALLREDUCE=peer)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29641",
         os.path.join(ROOT, "tests", "replica_worker.py")],
        env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "replica parity ok: world=2" in out.stdout
