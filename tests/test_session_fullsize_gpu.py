"""GPU: step-1 parity of the FULL-SIZE BASELINE graphs (configs[1], [2] and [3] per replica) --
the exact graphs bench.py times -- against the CPU oracle: loss, every gradient that
ApplyGradientDescent consumes, and the updated weights (tests/workloads.py::check_parity).
north_star tolerance: 1e-2 relative fp32 (relative Frobenius norm here; element-wise maxima of a
TF32 / bf16 path can flip single ReLU masks, see test_session_gpu.py::_lenet_step)."""
import pytest

import workloads as W
from simple_tensorflow_b200 import client

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["mlp", "lenet", "mlp_bf16"])
def test_full_size_training_step_vs_oracle(oracle, name):
    w = W.get(name)
    B = w.build(num_replicas=1, seed=1234, resident=False)
    with client.Session(B.tf.get_default_graph()) as sess:
        sess.run(B.tf.global_variables_initializer())
        p = W.check_parity(w, B, sess, oracle)
    print({k: v for k, v in p.items() if k != "what"})
    assert p["ok"], {k: v for k, v in p.items() if k != "what"}
    assert p["loss_rel_err"] < 1e-2 and p["weight_rel_err_max"] < 1e-2
    assert p["grad_rel_err_max_same_input_rounding"] < 1e-2
    # vs the exact fp32 oracle the ReLU-mask flips of a TF32 / bf16 forward bound the early layers
    assert p["grad_rel_err_max_vs_exact_fp32"] < (8e-2 if name == "mlp_bf16" else 4e-2)


def test_forwarding_never_overwrites_a_shared_or_fetched_tensor(oracle, rng):
    # Relu / BiasAdd / ReluGrad forward their input buffer in place when its
    # refcount is one; an entry that is fetched, or read by a second consumer, must keep its value
    import numpy as np
    from simple_tensorflow_b200 import ops as tf
    x = rng.randn(64, 96).astype(np.float32)
    b = rng.randn(96).astype(np.float32)
    tf.reset_default_graph()
    xp = tf.placeholder(tf.float32, [64, 96], "x")
    h = tf.bias_add(xp, tf.constant(b))          # single producer, two consumers below
    r = tf.relu(h)
    s = tf.add(h, r)                              # residual: add(x, relu(x))
    r2 = tf.relu(xp)                              # relu(placeholder) while the feed has other users
    with client.Session(tf.get_default_graph()) as sess:
        got_h, got_r = sess.run([h, r], {xp: x})  # pre-activation fetched together with its relu
        got_s = sess.run(s, {xp: x})
        got_r2, got_h2 = sess.run([r2, h], {xp: x})
    ref_h = x + b
    np.testing.assert_array_equal(got_h, ref_h)
    np.testing.assert_array_equal(got_r, np.maximum(ref_h, 0))
    np.testing.assert_array_equal(got_s, ref_h + np.maximum(ref_h, 0))
    np.testing.assert_array_equal(got_r2, np.maximum(x, 0))
    np.testing.assert_array_equal(got_h2, ref_h)


def test_tensors_may_outlive_their_session(rng):
    # TF_DeleteSession before TF_DeleteTensor is legal in the reference's C API
    import numpy as np
    from simple_tensorflow_b200 import ops as tf
    tf.reset_default_graph()
    xp = tf.placeholder(tf.float32, [8, 8], "x")
    y = tf.relu(xp)
    sess = client.Session(tf.get_default_graph())
    x = rng.randn(8, 8).astype(np.float32)
    out = sess.run(y, {xp: x}, as_host_tensors=True)
    staged = sess.stage(x)
    sess.close()
    del sess
    np.testing.assert_array_equal(np.array(out.numpy()), np.maximum(x, 0))
    del out
    staged.release()
