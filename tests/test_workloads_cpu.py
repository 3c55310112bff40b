"""CPU: the shared workload definitions (tests/workloads.py) -- bf16 helpers, the oracle chain of
the MLP against an independent float64 numpy implementation, the replica-average identity that
check_parity relies on for N > 1, and graph construction of the gradient fixes in ops.py."""
import numpy as np
import pytest

import workloads as W


def test_bf16_helpers_round_trip():
    a = np.array([1.0, 1.00390625, -3.14159, 65504.0, 1e-30, 0.1], np.float32)
    t = W.bf16_truncate(a)
    assert np.all(np.abs(t) <= np.abs(a))
    np.testing.assert_array_equal(W.bf16_from_bits(W.bf16_bits(t)), t)
    r = W.bf16_round(a)
    assert np.all(np.abs(r - a) <= np.abs(t - a) + 1e-45)            # nearest is never worse
    np.testing.assert_array_equal(W.bf16_round(r), r)                  # idempotent
    # ties go to even: 1 + 2^-8 is halfway between 1 and 1 + 2^-7
    assert W.bf16_round(np.array([1.0 + 2.0 ** -8], np.float32))[0] == 1.0
    assert W.bf16_round(np.array([1.0 + 3 * 2.0 ** -8], np.float32))[0] == np.float32(1.0 + 2.0 ** -6)


def _mlp_f64(x, labels, P, layers):
    acts = [x.astype(np.float64)]
    for i in range(layers):
        pre = acts[-1] @ P["W%d" % i].astype(np.float64) + P["b%d" % i].astype(np.float64)
        acts.append(np.maximum(pre, 0) if i < layers - 1 else pre)
    z = acts[-1] - acts[-1].max(1, keepdims=True)
    lse = np.log(np.exp(z).sum(1, keepdims=True))
    loss = float((labels * (lse - z)).sum(1).mean())
    g = (np.exp(z - lse) - labels) / x.shape[0]
    G = {}
    for i in reversed(range(layers)):
        G["b%d" % i] = g.sum(0)
        G["W%d" % i] = acts[i].T @ g
        if i > 0:
            g = (g @ P["W%d" % i].astype(np.float64).T) * (acts[i] > 0)
    return loss, G


@pytest.mark.parametrize("dtype,tol", [("f32", 2e-5), ("bf16", 6e-2)])
def test_mlp_oracle_chain_matches_float64_numpy(oracle, dtype, tol):
    w = W.MLP(dtype, batch=96, width=64, layers=3)
    x, labels = w.data(7)
    P = w.init_params()
    loss, G = w.reference(oracle, x, labels, P)
    ref_loss, ref_G = _mlp_f64(x, labels, P, 3)
    assert abs(loss - ref_loss) < max(tol, 1e-5) * abs(ref_loss)
    for n in G:
        assert W.rel_fro(G[n], ref_G[n]) < tol, n


def test_lenet_oracle_chain_shapes_and_descent(oracle):
    w = W.LeNet(batch=4)
    x, labels = w.data(3)
    P = w.init_params()
    loss0, G = w.reference(oracle, x, labels, P)
    assert set(G) == set(P) and all(G[n].shape == P[n].shape for n in P)
    # directional derivative along -grad matches the loss change of a small step
    eps = 1e-4
    Q = {n: P[n] - np.float32(eps) * G[n] for n in P}
    loss1, _ = w.reference(oracle, x, labels, Q)
    predicted = -eps * sum(float(np.sum(G[n].astype(np.float64) ** 2)) for n in P)
    assert loss1 < loss0
    assert abs((loss1 - loss0) - predicted) < 0.1 * abs(predicted)


def test_replica_average_of_gradients_is_the_global_batch_gradient(oracle):
    # what check_parity() uses for N > 1: mean over replicas of per-replica mean-loss gradients
    # equals the gradient of the mean loss over the concatenated batch
    w = W.MLP("f32", batch=32, width=48, layers=2)
    P = w.init_params()
    shards = [w.data(1234 + r) for r in range(2)]
    per = [w.reference(oracle, x, l, P)[1] for x, l in shards]
    xg = np.concatenate([s[0] for s in shards])
    lg = np.concatenate([s[1] for s in shards])
    glob = w.reference(oracle, xg, lg, P)[1]
    for n in glob:
        assert W.rel_fro(0.5 * (per[0][n] + per[1][n]), glob[n]) < 1e-5


def test_workload_flop_counts():
    assert W.get("mlp").flops_per_step == 8 * 2.0 * 4096 * 1024 * 1024
    le = W.get("lenet")
    assert abs(le.flops_per_step - (2 * 2.0 * 512 * 784 * 25 * 32 + 3 * 2.0 * 512 * 196 * 800 * 64 +
                                    3 * 2.0 * 512 * 3136 * 1024 + 3 * 2.0 * 512 * 1024 * 10)) < 1


def test_mean_gradient_is_materialised_for_non_xent_producers():
    # reduce_mean(relu(..)) must hand ReluGrad a gradient of the input's shape
    from simple_tensorflow_b200 import ops as tf
    tf.reset_default_graph()
    x = tf.placeholder(tf.float32, [4, 6], "x")
    v = tf.Variable(np.ones((6, 5), np.float32), name="v")
    loss = tf.reduce_mean(tf.relu(tf.matmul(x, v)))
    (g,) = tf.gradients(loss, [v])
    assert tf._shape(g) == (6, 5)
    relu_grads = [op for op in tf.get_default_graph().operations if op.type == "ReluGrad"]
    assert len(relu_grads) == 1 and tf._shape(relu_grads[0].inputs[0]) == (4, 5)


def test_weighted_xent_gradient_is_rejected_at_graph_construction():
    from simple_tensorflow_b200 import ops as tf
    tf.reset_default_graph()
    x = tf.placeholder(tf.float32, [4, 6], "x")
    lab = tf.placeholder(tf.float32, [4, 6], "l")
    v = tf.Variable(np.ones((6, 6), np.float32), name="v")
    per_example = tf.softmax_cross_entropy_with_logits(tf.matmul(x, v), lab)
    weights = tf.constant(np.arange(4, dtype=np.float32))
    loss = tf.reduce_mean(tf.multiply(per_example, weights))
    with pytest.raises(NotImplementedError):
        tf.gradients(loss, [v])


def test_tf32_input_rounding_moves_early_gradients_more_than_late_ones(oracle):
    # the effect KernelRounding exists for, on the CPU alone: the same oracle chain with its MatMul
    # operands truncated to TF32 differs from the exact chain by ~5e-4 in the last layer's gradient
    # but by far more in the first layer's (ReLU masks flipped by the forward rounding)
    w = W.MLP("f32", batch=512, width=256, layers=3)
    x, labels = w.data(11)
    P = w.init_params()
    _, G = w.reference(oracle, x, labels, P)
    _, Gk = w.reference(W.KernelRounding(oracle), x, labels, P)
    last, first = W.rel_fro(Gk["W2"], G["W2"]), W.rel_fro(Gk["W0"], G["W0"])
    assert last < 5e-3
    assert first > 2 * last
    assert first < 5e-2
