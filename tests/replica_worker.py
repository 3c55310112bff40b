"""Worker of tests/test_replica_gpu.py: one rank of a world_size-N replica data-parallel run.

Every rank builds the same MLP graph (identical initial weights), trains STEPS steps on its own
shard of a global batch through Session.run with staged feeds, and rank 0 checks the resulting
variables against the CPU oracle trained on the WHOLE batch: the average over replicas of
per-shard mean gradients is the full-batch mean gradient, so the two must agree to the
floating-point bar.  Launched by torchrun; not collected by pytest.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402  (rendezvous only)
import torch.distributed as dist  # noqa: E402

from simple_tensorflow_b200 import _lib, client, replica  # noqa: E402
from simple_tensorflow_b200 import ops as tf  # noqa: E402

STEPS, B, D, LR = 2, 512, 256, 0.5


def main():
    bucket = os.environ.get("REPLICA_TEST_BUCKET", "default")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.load()
    comm = replica.init_nccl_comm(L, rank, world, local)

    rng = np.random.RandomState(77)
    x = rng.uniform(-1, 1, (B, D)).astype(np.float32)
    labels = np.eye(D, dtype=np.float32)[rng.randint(0, D, B)]
    ws = [(rng.randn(D, D) / np.sqrt(D)).astype(np.float32) for _ in range(3)]
    bs = [np.full(D, 0.1, np.float32) for _ in range(3)]
    lo, hi = replica.shard_batch(B, world, rank)

    tf.reset_default_graph()
    xp, lp = tf.placeholder(tf.float32, [hi - lo, D]), tf.placeholder(tf.float32, [hi - lo, D])
    Ws = [tf.Variable(w, name="W%d" % i) for i, w in enumerate(ws)]
    Bs = [tf.Variable(b, name="b%d" % i) for i, b in enumerate(bs)]
    h = xp
    for i in range(3):
        h = tf.bias_add(tf.matmul(h, Ws[i]), Bs[i])
        if i < 2:
            h = tf.relu(h)
    loss = tf.reduce_mean(tf.softmax_cross_entropy_with_logits(h, lp))
    kw = {} if bucket == "default" else {"bucket_bytes": None if bucket == "none" else int(bucket)}
    train = tf.GradientDescentOptimizer(LR).minimize(loss, num_replicas=world, **kw)
    n_collectives = sum(1 for op in tf.get_default_graph().operations if op.type == "B200AllReduceN")

    sess = client.Session(tf.get_default_graph(), gpu=local, collective_comm=comm, num_replicas=world)
    sess.run(tf.global_variables_initializer())
    hx, hl = client.HostTensor.from_numpy(x[lo:hi]), client.HostTensor.from_numpy(labels[lo:hi])
    losses = []
    staged = (sess.stage(hx), sess.stage(hl))
    for step in range(STEPS):
        nxt = (sess.stage(hx), sess.stage(hl)) if step + 1 < STEPS else None
        losses.append(float(sess.run([loss, train], {xp: staged[0], lp: staged[1]})[0]))
        staged = nxt
    got = sess.run([v.ref for v in Ws] + [v.ref for v in Bs])

    # every replica must hold the same variables (the update used identical averaged gradients)
    for g in got:
        t = torch.from_numpy(np.array(g)).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), "replicas diverged"
    mean_loss = torch.tensor([losses[0]], device="cuda")
    dist.all_reduce(mean_loss)
    if rank == 0:
        import oracle_bind as oracle
        from test_session_gpu import _mlp_reference
        oracle.lib()
        rw, rb, ref_loss0 = ws, bs, None
        for _ in range(STEPS):
            l, rw, rb = _mlp_reference(oracle, x, labels, rw, rb, LR)
            ref_loss0 = l if ref_loss0 is None else ref_loss0
        assert abs(mean_loss.item() / world - ref_loss0) < 1e-2 * abs(ref_loss0), (mean_loss, ref_loss0)
        for g, r in zip(got, rw + rb):
            err = np.abs(g - r).max() / np.abs(r).max()
            assert err < 1e-2, err
        print("replica parity ok: world=%d collectives=%d losses=%s" % (world, n_collectives, losses))
    dist.barrier()
    sess.close()
    _lib.check(L.b200_nccl_comm_destroy(comm))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
