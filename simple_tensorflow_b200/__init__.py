"""simple_tensorflow_b200 -- B200-native op-kernel layer for the stripped TensorFlow 1.0 core of
DengZhuangSouthRd/simple_tensorflow.

Layers (bottom up):
  csrc/*.cu                 hand-written sm_100a kernels behind the C ABI include/b200_ops.h
                            -> lib/libb200tf.so                      (binding: _lib.py)
  csrc/tensorflow/...       C++ mirror of the reference's plugin surface (REGISTER_OP /
                            REGISTER_KERNEL_BUILDER / OpKernelContext), B200 GPU device,
                            DirectSession-contract executor, C API
                            -> lib/libb200tf_framework.so            (binding: client.py)
  ops.py                    Python op constructors + gradients + SGD (what the reference's
                            python/ops and python/training do for these ops)
  compat.py                 the same under the reference's names (tf.nn.*, tf.train.*, tf.Session)
Nothing here falls back to CPU or PyTorch math: without the built libraries imports fail loudly.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "client", "ops", "compat"]
