"""`import simple_tensorflow_b200.compat as tf`: the names a TensorFlow-1.0 script uses for the
hot path, in the reference's module layout (python/ops/nn.py, python/training/, python/client/),
mapped onto ops.py / client.py.  Only what exists underneath is exposed -- nothing is emulated.

    import simple_tensorflow_b200.compat as tf
    x = tf.placeholder(tf.float32, [4096, 1024])                    # static shapes only
    h = tf.nn.relu(tf.nn.bias_add(tf.matmul(x, W), b))
    loss = tf.reduce_mean(tf.nn.softmax_cross_entropy_with_logits(logits=h, labels=y))
    train = tf.train.GradientDescentOptimizer(0.1).minimize(loss)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        sess.run([loss, train], {x: ..., y: ...})
"""
from . import client as _client
from . import ops as _ops
from .client import Graph, float32, int32, int64, bfloat16  # noqa: F401
from .ops import (Variable, add, add_n, argmax, batch_matmul, cast, constant,  # noqa: F401
                  get_default_graph, global_variables_initializer, gradients, group, identity,
                  import_graph_def, matmul, multiply, placeholder, reduce_mean, reset_default_graph,
                  reshape)


class Session(_client.Session):
    """session.py Session: the graph argument defaults to the default graph."""

    def __init__(self, graph=None, **kw):
        super().__init__(graph if graph is not None else get_default_graph(), **kw)


class _NN:
    """tensorflow.python.ops.nn: reference keyword names (logits= / labels=, value= / ksize=)."""
    relu = staticmethod(_ops.relu)
    softmax = staticmethod(_ops.softmax)
    log_softmax = staticmethod(_ops.log_softmax)
    bias_add = staticmethod(_ops.bias_add)
    conv2d = staticmethod(_ops.conv2d)
    max_pool = staticmethod(_ops.max_pool)

    @staticmethod
    def softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        # nn_ops.py softmax_cross_entropy_with_logits(_sentinel, labels, logits): named arguments only
        if _sentinel is not None or labels is None or logits is None:
            raise ValueError("Only call `softmax_cross_entropy_with_logits` with named arguments "
                             "(labels=..., logits=..., ...)")
        return _ops.softmax_cross_entropy_with_logits(logits, labels, name=name)


class _Train:
    GradientDescentOptimizer = _ops.GradientDescentOptimizer


nn = _NN()
train = _Train()
