"""GPU parity: every hot-path op, called through the C ABI (include/b200_ops.h), against the
CPU oracle on identical inputs, the reference's golden vectors, and size-independent
properties at BASELINE.json's full sizes.

Bars (BASELINE.json north_star): <= 1e-2 relative fp32 for floating point -- measured as
max|got - ref| / max|ref| (abi_util.rel_err); bit-exact for Cast / ArgMax / MaxPool (pure
selection) / index outputs.  bf16 parity (SURVEY 7 "hard parts"): inputs are truncated to bf16
exactly as the reference's Cast does, the fp32 oracle runs on the truncated values, tolerance
1e-2 relative (bf16 output rounding alone is 2^-9 = 2e-3).
"""
import numpy as np
import pytest

import abi_util as au
import golden_util as gu

pytestmark = pytest.mark.gpu

TOL = 1e-2          # north_star: within 1e-2 relative fp32
TOL_TF32 = 3e-3     # what single-pass TF32 actually achieves on these shapes (tighter guard)


@pytest.fixture(scope="module", autouse=True)
def _device():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    L = au.lib()
    assert L.b200_device_count() >= 1
    before = L.b200_launch_count()
    yield
    assert L.b200_launch_count() > before, "no kernels of libb200tf.so were launched"


# =============================================================================== MatMul
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 384, 512), (200, 136, 72), (129, 4, 36),
                                   (1000, 520, 260), (4096, 1024, 1024)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_matmul_tcgen05_vs_oracle(oracle, rng, m, n, k, ta, tb):
    a = rng.uniform(-1, 1, (k, m) if ta else (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, k) if tb else (k, n)).astype(np.float32)
    got = au.matmul(a, b, ta, tb)
    ref = oracle.matmul(a, b, ta, tb)
    assert not np.isnan(got).any()
    assert au.rel_err(got, ref) < TOL_TF32


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (3, 5, 1), (5, 3, 5), (1, 5, 3), (37, 53, 71)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_matmul_small_unaligned_exact_fp32(oracle, rng, m, n, k, ta, tb):
    # the reference's own sizes/tolerance: matmul_op_test.py:209-253 ({1,3,5}^3, N(-5,5), 1e-5)
    a = rng.normal(-5, 5, (k, m) if ta else (m, k)).astype(np.float32)
    b = rng.normal(-5, 5, (n, k) if tb else (k, n)).astype(np.float32)
    np.testing.assert_allclose(au.matmul(a, b, ta, tb), oracle.matmul(a, b, ta, tb), rtol=1e-5,
                               atol=1e-3)


@pytest.mark.parametrize("m,n,k", [(512, 10, 1024), (1024, 10, 512), (100, 7, 65), (64, 30, 64),
                                   (333, 17, 1000), (65, 1, 130)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_matmul_skinny_n(oracle, rng, m, n, k, ta, tb):
    # N <= 32 (LeNet's 10-class layer and its weight gradient): the K-split CUDA-core kernel, both
    # thread maps (A stored [M, K] / [K, M]); IEEE fp32 FMAs, so the bar is summation order only
    a = rng.uniform(-1, 1, (k, m) if ta else (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, k) if tb else (k, n)).astype(np.float32)
    got = au.matmul(a, b, ta, tb)
    ref = oracle.matmul(a, b, ta, tb)
    # B stored [K, N] with N % 4 != 0 cannot be a TMA operand: that is the CUDA-core path (exact
    # fp32); B stored [N, K] may take the tensor path (TF32 tolerance)
    assert au.rel_err(got, ref) < (TOL_TF32 if tb else 1e-5)
    np.testing.assert_array_equal(got, au.matmul(a, b, ta, tb))


@pytest.mark.parametrize("shape", [(3, 1000, 200, 1), (5, 1, 200, 1000), (2, 70, 64, 3), (4, 2, 100, 333),
                                   (1, 9, 129, 640)])
@pytest.mark.parametrize("adj_x", [False, True])
@pytest.mark.parametrize("adj_y", [False, True])
def test_batch_matmul_vector_shapes(oracle, rng, shape, adj_x, adj_y):
    # matrix-vector / vector-matrix products (batch_matmul_op_test.cc:107-132): the K-split kernel,
    # batched; skinny-M runs as the transposed problem
    batch, m, k, n = shape
    x = rng.uniform(-1, 1, (batch, k, m) if adj_x else (batch, m, k)).astype(np.float32)
    y = rng.uniform(-1, 1, (batch, n, k) if adj_y else (batch, k, n)).astype(np.float32)
    got = au.batch_matmul(x, y, adj_x, adj_y)
    assert au.rel_err(got, oracle.batch_matmul(x, y, adj_x, adj_y)) < TOL_TF32


def test_matmul_integer_inputs_exact(oracle, rng):
    # integers <= 2048 are exact in tf32 and their sums are exact in the fp32 accumulator
    a = rng.randint(-8, 9, (256, 192)).astype(np.float32)
    b = rng.randint(-8, 9, (192, 320)).astype(np.float32)
    np.testing.assert_array_equal(au.matmul(a, b), oracle.matmul(a, b))


def test_matmul_direct_session_known_answer():
    # core/common_runtime/direct_session_test.cc:88-108: [[3,2],[-1,0]] * [[1],[1]] = [[5],[-1]]
    got = au.matmul(np.array([[3, 2], [-1, 0]], np.float32), np.array([[1], [1]], np.float32))
    np.testing.assert_array_equal(got.ravel(), [5.0, -1.0])


def test_matmul_precision_mode_simt(oracle, rng):
    L = au.lib()
    a = rng.normal(-5, 5, (256, 256)).astype(np.float32)
    b = rng.normal(-5, 5, (256, 256)).astype(np.float32)
    assert L.b200_set_matmul_precision(1) == 0
    try:
        np.testing.assert_allclose(au.matmul(a, b), oracle.matmul(a, b), rtol=1e-5, atol=1e-2)
    finally:
        assert L.b200_set_matmul_precision(0) == 0
    assert L.b200_set_matmul_precision(7) == 3  # INVALID_ARGUMENT


def test_matmul_splitk_matches_unsplit(oracle, rng):
    # dW shape of the MLP: 1024x1024 output, K = 4096 -> split-K with scratch, ordered reduction
    a = rng.uniform(-1, 1, (4096, 1024)).astype(np.float32)
    b = rng.uniform(-1, 1, (4096, 1024)).astype(np.float32)
    ref = oracle.matmul(a, b, True, False)
    split = au.matmul(a, b, True, False, use_workspace=True)
    nosplit = au.matmul(a, b, True, False, use_workspace=False)
    assert au.rel_err(split, ref) < TOL_TF32 and au.rel_err(nosplit, ref) < TOL_TF32
    # run-to-run determinism of the ordered reduction
    np.testing.assert_array_equal(split, au.matmul(a, b, True, False, use_workspace=True))


@pytest.mark.parametrize("m,n,k", [(4096, 1024, 1024), (512, 384, 256), (200, 136, 72),
                                   (128, 64, 96), (37, 53, 71)])
def test_fused_matmul_bias_relu_and_relu_grad(oracle, rng, m, n, k):
    a = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    bias = rng.uniform(-1, 1, n).astype(np.float32)
    plain = oracle.matmul(a, b)
    tol = TOL_TF32 * np.abs(plain).max()
    got = au.fused_matmul(a, b, bias=bias)
    assert np.abs(got - oracle.bias_add(plain, bias)).max() < tol
    got = au.fused_matmul(a, b, bias=bias, relu=True)
    ref = oracle.relu(oracle.bias_add(plain, bias))
    assert np.abs(got - ref).max() < tol and (got >= 0).all()
    feat = rng.uniform(-1, 1, (m, n)).astype(np.float32)
    feat[::3] = 0.0
    got = au.fused_matmul(a, b, features=feat)
    assert np.abs(got - oracle.relu_grad(plain, feat)).max() < tol
    np.testing.assert_array_equal(got[feat <= 0], 0.0)
    # transposed operands (the dX = dY * W^T shape of the backward pass)
    bt = np.ascontiguousarray(b.T)
    got = au.fused_matmul(a, bt, tb=True, features=feat)
    assert np.abs(got - oracle.relu_grad(plain, feat)).max() < tol


def test_fused_matmul_bf16(oracle, rng):
    m, n, k = 512, 256, 384
    a = oracle.truncate_to_bf16(rng.uniform(-1, 1, (m, k)).astype(np.float32))
    b = oracle.truncate_to_bf16(rng.uniform(-1, 1, (k, n)).astype(np.float32))
    bias = oracle.truncate_to_bf16(rng.uniform(-1, 1, n).astype(np.float32))
    ref = oracle.relu(oracle.bias_add(oracle.matmul(a, b), bias))
    assert au.rel_err(au.fused_matmul(a, b, bias=bias, relu=True, bf16=True), ref) < TOL


def test_matmul_rejects_bad_arguments():
    L = au.lib()
    assert L.b200_matmul(1, 16, 16, 16, 0, 4, 4, 0, 0, None, 0, None) == 3   # m == 0
    assert L.b200_matmul(2, 16, 16, 16, 4, 4, 4, 0, 0, None, 0, None) == 12  # DT_DOUBLE
    assert b"dtype" in L.b200_last_error()


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_matmul_bf16(oracle, rng, ta, tb):
    m, n, k = 512, 384, 640
    a = oracle.truncate_to_bf16(rng.uniform(-1, 1, (k, m) if ta else (m, k)).astype(np.float32))
    b = oracle.truncate_to_bf16(rng.uniform(-1, 1, (n, k) if tb else (k, n)).astype(np.float32))
    got = au.matmul(a, b, ta, tb, bf16=True)
    assert au.rel_err(got, oracle.matmul(a, b, ta, tb)) < TOL


@pytest.mark.parametrize("adj_x", [False, True])
@pytest.mark.parametrize("adj_y", [False, True])
@pytest.mark.parametrize("shape", [(3, 5, 7, 4), (4, 256, 128, 192), (2, 130, 100, 68)])
def test_batch_matmul(oracle, rng, adj_x, adj_y, shape):
    batch, m, k, n = shape
    x = rng.uniform(-1, 1, (batch, k, m) if adj_x else (batch, m, k)).astype(np.float32)
    y = rng.uniform(-1, 1, (batch, n, k) if adj_y else (batch, k, n)).astype(np.float32)
    got = au.batch_matmul(x, y, adj_x, adj_y)
    assert au.rel_err(got, oracle.batch_matmul(x, y, adj_x, adj_y)) < TOL_TF32


# =============================================================================== BiasAdd(+Grad)
@pytest.mark.parametrize("shape", [(4096, 1024), (512, 28, 28, 32), (512, 10), (7, 3, 5), (3, 1),
                                   (64, 14, 14, 64), (5, 1030)])
def test_bias_add(oracle, rng, shape):
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    b = rng.uniform(-1, 1, shape[-1]).astype(np.float32)
    np.testing.assert_array_equal(au.bias_add(x, b), oracle.bias_add(x, b))  # one fp32 add: exact


@pytest.mark.parametrize("shape", [(4096, 1024), (512, 28, 28, 32), (512, 10), (7, 3, 5), (3, 1),
                                   (64, 14, 14, 64), (1000, 1030), (1, 8)])
def test_bias_add_grad(oracle, rng, shape):
    g = rng.uniform(-1, 1, shape).astype(np.float32)
    got = au.bias_add_grad(g)
    ref = g.reshape(-1, shape[-1]).astype(np.float64).sum(0)
    assert au.rel_err(got, ref) < 1e-5
    assert au.rel_err(oracle.bias_add_grad(g), ref) < 1e-3  # oracle sums sequentially in fp32
    np.testing.assert_array_equal(got, au.bias_add_grad(g))  # deterministic (no atomics)


def test_bias_add_bf16(oracle, rng):
    x = oracle.truncate_to_bf16(rng.uniform(-1, 1, (256, 264)).astype(np.float32))
    b = oracle.truncate_to_bf16(rng.uniform(-1, 1, 264).astype(np.float32))
    assert au.rel_err(au.bias_add(x, b, bf16=True), oracle.bias_add(x, b)) < TOL
    assert au.rel_err(au.bias_add_grad(x, bf16=True), oracle.bias_add_grad(x)) < TOL


NCHW_SHAPES = [(4, 8, 5, 7), (3, 16, 8, 8), (2, 3, 80, 80), (5, 6, 4), (2, 3, 4, 12, 20),
               (512, 64, 14, 14), (1, 1, 1, 1)]


@pytest.mark.parametrize("shape", NCHW_SHAPES)
def test_bias_add_nchw_native(rng, shape):
    # bias_op_gpu.cu.cc:56-63: out[.., c, h, w] = in + bias[c]; channel = dims - 3
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    b = rng.uniform(-1, 1, shape[-3]).astype(np.float32)
    np.testing.assert_array_equal(au.bias_add_nchw(x, b), x + b[:, None, None])


@pytest.mark.parametrize("shape", NCHW_SHAPES)
def test_bias_add_grad_nchw_native(rng, shape):
    g = rng.uniform(-1, 1, shape).astype(np.float32)
    got = au.bias_add_grad_nchw(g)
    axes = tuple(i for i in range(len(shape)) if i != len(shape) - 3)
    ref = g.astype(np.float64).sum(axes)
    assert got.shape == (shape[-3],)
    assert au.rel_err(got, ref) < 1e-5
    np.testing.assert_array_equal(got, au.bias_add_grad_nchw(g))  # ordered, no atomics


def test_bias_nchw_bf16(oracle, rng):
    x = oracle.truncate_to_bf16(rng.uniform(-1, 1, (6, 24, 9, 16)).astype(np.float32))
    b = oracle.truncate_to_bf16(rng.uniform(-1, 1, 24).astype(np.float32))
    assert au.rel_err(au.bias_add_nchw(x, b, bf16=True), x + b[:, None, None]) < TOL
    assert au.rel_err(au.bias_add_grad_nchw(x, bf16=True), x.astype(np.float64).sum((0, 2, 3))) < TOL
    y = oracle.truncate_to_bf16(rng.uniform(-1, 1, (2, 3, 5, 7)).astype(np.float32))  # scalar path
    assert au.rel_err(au.bias_add_grad_nchw(y, bf16=True), y.astype(np.float64).sum((0, 2, 3))) < TOL


# =============================================================================== Relu(+Grad)
@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 * 1024 + 3])
def test_relu_and_grad(oracle, rng, n):
    x = rng.uniform(-1, 1, n).astype(np.float32)
    x[::7] = 0.0
    g = rng.uniform(-1, 1, n).astype(np.float32)
    np.testing.assert_array_equal(au.relu(x), oracle.relu(x))
    np.testing.assert_array_equal(au.relu_grad(g, x), oracle.relu_grad(g, x))


def test_relu_known_values():
    # relu_op_test.py:37-62
    x = np.array([[-0.9, 0.7, -0.5, 0.3, -0.1], [0.1, -0.3, 0.5, -0.7, 0.9]], np.float32)
    np.testing.assert_array_equal(au.relu(x), np.maximum(x, 0))


def test_relu_bf16(oracle, rng):
    x = oracle.truncate_to_bf16(rng.uniform(-1, 1, 4099).astype(np.float32))
    g = oracle.truncate_to_bf16(rng.uniform(-1, 1, 4099).astype(np.float32))
    np.testing.assert_array_equal(au.relu(x, bf16=True), oracle.relu(x))
    np.testing.assert_array_equal(au.relu_grad(g, x, bf16=True), oracle.relu_grad(g, x))


# =============================================================================== Softmax / xent
@pytest.mark.parametrize("shape", [(2, 4), (512, 10), (4096, 1024), (33, 1000), (7, 1), (5, 2048),
                                   (3, 5000), (64, 132)])
@pytest.mark.parametrize("log", [False, True])
def test_softmax(oracle, rng, shape, log):
    x = (rng.randn(*shape) * 3).astype(np.float32)
    got = au.softmax(x, log)
    ref = oracle.softmax(x, log)
    # expf/logf differ by an ulp or two between libm and the GPU; the bar is 1e-2 relative
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
    if not log:
        np.testing.assert_allclose(got.sum(1), np.ones(shape[0]), rtol=1e-5)


def test_softmax_known_answers_and_overflow():
    # softmax_op_test.py:76-117
    f = np.array([[1., 1., 1., 1.], [1., 2., 3., 4.]], np.float32)
    np.testing.assert_allclose(au.softmax(f), [[0.25] * 4,
                                               [0.0320586, 0.08714432, 0.23688282, 0.64391426]],
                               rtol=1e-5, atol=1e-5)
    mx = np.finfo(np.float32).max
    f = np.array([[1., 1., 1., 1.], [mx, 1., 2., 3.]], np.float32)
    np.testing.assert_allclose(au.softmax(f, log=True), [[-1.386294] * 4, [0, -mx, -mx, -mx]],
                               rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 4), (512, 10), (4096, 1024), (9, 1500)])
def test_softmax_xent(oracle, rng, shape):
    x = (rng.randn(*shape) * 2).astype(np.float32)
    labels = np.zeros(shape, np.float32)
    labels[np.arange(shape[0]), rng.randint(0, shape[1], shape[0])] = 1.0
    loss, bp = au.softmax_xent(x, labels)
    rloss, rbp = oracle.softmax_xent(x, labels)
    np.testing.assert_allclose(loss, rloss, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(bp, rbp, rtol=1e-4, atol=1e-6)


def test_xent_known_answers():
    # xent_op_test.py:96-131
    f = np.array([[1., 1., 1., 1.], [1., 2., 3., 4.]], np.float32)
    l = np.array([[0., 0., 0., 1.], [0., .5, .5, 0.]], np.float32)
    loss, bp = au.softmax_xent(f, l)
    np.testing.assert_allclose(bp, [[0.25, 0.25, 0.25, -0.75], [0.0321, -0.4129, -0.2632, 0.6439]],
                               rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(loss, [1.3862, 1.9401], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("shape", [(4096, 1024), (33, 264), (17, 512), (64, 8), (9, 1000), (5, 1030),
                                   (3, 2048)])
def test_softmax_xent_bf16(oracle, rng, shape):
    # vectorised (cols % 8 == 0, <= 1024), row-per-warp scalar and block kernels: fp32 math, one
    # rounding to bf16 per output
    x = oracle.truncate_to_bf16((rng.randn(*shape) * 2).astype(np.float32))
    labels = np.zeros(shape, np.float32)
    labels[np.arange(shape[0]), rng.randint(0, shape[1], shape[0])] = 1.0
    loss, bp = au.softmax_xent(x, labels, bf16=True)
    rloss, rbp = oracle.softmax_xent(x, labels)
    assert au.rel_err(loss, rloss) < TOL
    assert np.max(np.abs(bp - rbp)) < 1e-2  # |backprop| <= 1; bf16 keeps 8 bits


def test_softmax_bf16(oracle, rng):
    x = oracle.truncate_to_bf16((rng.randn(64, 1024) * 2).astype(np.float32))
    assert np.max(np.abs(au.softmax(x, bf16=True) - oracle.softmax(x))) < 1e-2 * oracle.softmax(x).max()


# =============================================================================== MaxPool(+Grad)
@pytest.mark.parametrize("case", gu.pool_cases("max_pool"), ids=gu.case_id)
def test_max_pool_golden(oracle, case):
    x = gu.iota(case["input_sizes"])
    got = au.max_pool(x, case["ksize"][1:3], case["strides"][1:3], case["padding"], oracle)
    np.testing.assert_array_equal(got.ravel(), np.asarray(case["expected"], np.float32))


@pytest.mark.parametrize("case", gu.pool_cases("max_pool_grad"), ids=gu.case_id)
def test_max_pool_grad_direct_golden(oracle, case):
    x = np.asarray(case["input_data"], np.float32).reshape(case["input_sizes"])
    g = np.asarray(case["output_backprop"], np.float32).reshape(case["output_sizes"])
    got = au.max_pool_grad(x, g, [case["window_rows"], case["window_cols"]],
                           [case["row_stride"], case["col_stride"]], case["padding"], oracle)
    np.testing.assert_array_equal(got.ravel(),
                                  np.asarray(case["expected_input_backprop"], np.float32))


@pytest.mark.parametrize("shape,ksize,strides,padding", [
    ((512, 28, 28, 32), (2, 2), (2, 2), "SAME"), ((64, 14, 14, 64), (2, 2), (2, 2), "SAME"),
    ((4, 17, 19, 6), (3, 3), (2, 2), "VALID"), ((3, 9, 9, 5), (3, 2), (1, 2), "SAME"),
    ((2, 8, 8, 8), (3, 3), (2, 2), "SAME"), ((2, 7, 7, 3), (2, 2), (3, 3), "VALID"),
])
def test_max_pool_and_grad_vs_oracle(oracle, rng, shape, ksize, strides, padding):
    x = rng.randint(0, 6, shape).astype(np.float32)  # many ties -> exercises first-max rule
    y = au.max_pool(x, ksize, strides, padding, oracle)
    np.testing.assert_array_equal(y, oracle.max_pool(x, ksize, strides, padding))
    g = rng.randint(1, 5, y.shape).astype(np.float32)  # small ints: sums exact in any order
    np.testing.assert_array_equal(au.max_pool_grad(x, g, ksize, strides, padding, oracle),
                                  oracle.max_pool_grad(x, g, ksize, strides, padding))


def test_max_pool_bf16(oracle, rng):
    x = oracle.truncate_to_bf16(rng.randn(8, 14, 14, 64).astype(np.float32))
    np.testing.assert_array_equal(au.max_pool(x, (2, 2), (2, 2), "SAME", oracle, bf16=True),
                                  oracle.max_pool(x, (2, 2), (2, 2), "SAME"))


# =============================================================================== Cast / ArgMax
def test_cast_bfloat16_bit_exact(oracle, rng):
    x = np.concatenate([rng.randn(100003).astype(np.float32) * 1e3,
                        np.array([0.0, -0.0, np.inf, -np.inf, 1e-40, 3.3895e38], np.float32)])
    b = au.cast(x, np.float32, au.BF16)
    np.testing.assert_array_equal(b, oracle.cast_f32_to_bf16(x))  # truncation, not rounding
    back = au.cast(b, au.BF16, np.float32)
    np.testing.assert_array_equal(back.view(np.uint32), oracle.cast_bf16_to_f32(b).view(np.uint32))
    x8 = rng.randn(4096).astype(np.float32)  # the vectorised (multiple-of-8) path
    np.testing.assert_array_equal(au.cast(x8, np.float32, au.BF16), oracle.cast_f32_to_bf16(x8))
    b8 = oracle.cast_f32_to_bf16(x8)
    np.testing.assert_array_equal(au.cast(b8, au.BF16, np.float32), oracle.cast_bf16_to_f32(b8))


def test_cast_numeric_bit_exact(oracle, rng):
    f = (rng.randn(1001) * 1000).astype(np.float32)
    np.testing.assert_array_equal(au.cast(f, np.float32, np.int32), oracle.cast(f, np.int32))
    np.testing.assert_array_equal(au.cast(f, np.float32, np.int64), oracle.cast(f, np.int64))
    i = rng.randint(-2 ** 31, 2 ** 31 - 1, 1001).astype(np.int32)
    np.testing.assert_array_equal(au.cast(i, np.int32, np.float32), oracle.cast(i, np.float32))
    np.testing.assert_array_equal(au.cast(i, np.int32, np.int64), oracle.cast(i, np.int64))
    j = rng.randint(-2 ** 40, 2 ** 40, 1001).astype(np.int64)
    np.testing.assert_array_equal(au.cast(j, np.int64, np.float32), oracle.cast(j, np.float32))
    np.testing.assert_array_equal(au.cast(j, np.int64, np.int32), oracle.cast(j, np.int32))


def test_argmax_every_axis_bit_exact(oracle, rng):
    # argmax_op_test.py:28-68
    x = rng.randn(3, 2, 4, 5, 6).astype(np.float32)
    for axis in range(-5, 5):
        got = au.argmax(x, axis)
        np.testing.assert_array_equal(got, np.argmax(x, axis=axis))
        np.testing.assert_array_equal(got, oracle.argmax(x, axis))
    xi = rng.randint(0, 3, (4, 7, 3)).astype(np.int32)
    for axis in range(3):
        np.testing.assert_array_equal(au.argmax(xi, axis), oracle.argmax(xi, axis))


@pytest.mark.parametrize("shape", [(4096, 1024), (512, 10), (100, 64), (33, 3000)])
def test_argmax_rows_with_ties(oracle, rng, shape):
    x = rng.randint(0, 50, shape).astype(np.float32)  # ties: lowest index must win
    np.testing.assert_array_equal(au.argmax(x, 1), oracle.argmax(x, 1))
    np.testing.assert_array_equal(au.argmax(x, 0), oracle.argmax(x, 0))
    const = np.full((70, 200), 3.0, np.float32)
    np.testing.assert_array_equal(au.argmax(const, 1), np.zeros(70, np.int64))


# =============================================================================== Conv2D family
@pytest.mark.parametrize("case", gu.conv_cases("conv2d"), ids=gu.case_id)
def test_conv2d_golden(oracle, case):
    x = gu.iota(case["tensor_in_sizes"])
    f = gu.iota(case["filter_in_sizes"])
    got = au.conv2d(x, f, case["strides"], case["padding"], oracle)
    # inputs are integers <= 36: exact in tf32; sums exact in fp32 (conv_ops_test.py tol 1e-5)
    np.testing.assert_allclose(got.ravel(), np.asarray(case["expected"], np.float32), rtol=1e-5,
                               atol=1e-5)


@pytest.mark.parametrize("case", gu.conv_cases("conv2d_backprop_input"), ids=gu.case_id)
def test_conv2d_backprop_input_golden(oracle, case):
    f = gu.iota(case["filter_sizes"])
    dy = gu.iota(case["output_sizes"])
    got = au.conv2d_backprop_input(case["input_sizes"], f, dy, case["strides"], case["padding"],
                                   oracle)
    np.testing.assert_allclose(got.ravel(), np.asarray(case["expected"], np.float32), rtol=1e-5,
                               atol=1e-4)


@pytest.mark.parametrize("case", gu.conv_cases("conv2d_backprop_filter"), ids=gu.case_id)
def test_conv2d_backprop_filter_golden(oracle, case):
    x = gu.iota(case["input_sizes"])
    dy = gu.iota(case["output_sizes"])
    got = au.conv2d_backprop_filter(x, case["filter_sizes"], dy, case["strides"], case["padding"],
                                    oracle)
    np.testing.assert_allclose(got.ravel(), np.asarray(case["expected"], np.float32), rtol=1e-5,
                               atol=1e-4)


CONV_SHAPES = [
    # LeNet (BASELINE config 3) at reduced batch, then ragged / strided / pointwise cases
    ((16, 28, 28, 1), (5, 5, 1, 32), (1, 1), "SAME"),
    ((16, 14, 14, 32), (5, 5, 32, 64), (1, 1), "SAME"),
    ((4, 17, 13, 8), (3, 3, 8, 16), (2, 2), "SAME"),
    ((2, 9, 8, 3), (3, 2, 3, 4), (2, 1), "VALID"),
    ((3, 12, 12, 16), (1, 1, 16, 24), (1, 1), "VALID"),
    ((2, 7, 7, 4), (7, 7, 4, 8), (1, 1), "VALID"),
    ((2, 10, 10, 6), (3, 3, 6, 10), (1, 1), "SAME"),
    # channel counts that take the implicit-GEMM path (TMA im2col): strides, odd sizes, SAME pads
    ((2, 15, 13, 32), (3, 3, 32, 16), (2, 2), "SAME"),
    ((3, 9, 11, 64), (3, 2, 64, 32), (1, 2), "VALID"),
    ((2, 8, 8, 96), (1, 3, 96, 40), (1, 1), "SAME"),
    ((5, 7, 7, 32), (7, 7, 32, 8), (1, 1), "SAME"),
    ((1, 33, 35, 32), (5, 5, 32, 64), (3, 2), "SAME"),
    # first layers (C_in <= 4): the direct CUDA-core kernels (forward; filter gradient when K % 32 == 0)
    ((2, 12, 12, 3), (3, 3, 3, 64), (1, 1), "SAME"),
    ((3, 11, 9, 1), (5, 5, 1, 32), (2, 2), "SAME"),
    ((2, 8, 8, 4), (3, 3, 4, 96), (1, 1), "VALID"),     # 36 taps: two tap blocks
    ((2, 6, 7, 2), (7, 5, 2, 32), (1, 2), "SAME"),      # 70 taps: three tap blocks
    # one input channel, unit stride: the lane-per-filter sliding-window kernels (conv_c1_*)
    ((5, 13, 11, 1), (5, 5, 1, 64), (1, 1), "SAME"),    # two filter blocks, widths not multiples of S
    ((3, 9, 16, 1), (3, 3, 1, 32), (1, 1), "VALID"),
    ((2, 6, 5, 1), (5, 5, 1, 32), (1, 1), "VALID"),     # 2 x 1 output pixels
    ((9, 7, 30, 1), (3, 3, 1, 96), (1, 1), "SAME"),     # fewer rows than warps
]


@pytest.mark.parametrize("shape,fshape,strides,padding", CONV_SHAPES)
def test_conv_family_vs_oracle(oracle, rng, shape, fshape, strides, padding):
    # conv_ops_test.py:246-247 uses np.random.rand inputs for CPU-vs-GPU compares
    x = rng.rand(*shape).astype(np.float32)
    f = rng.rand(*fshape).astype(np.float32)
    y_ref = oracle.conv2d(x, f, strides, padding)
    assert au.rel_err(au.conv2d(x, f, strides, padding, oracle), y_ref) < TOL_TF32
    dy = rng.rand(*y_ref.shape).astype(np.float32)
    dx = au.conv2d_backprop_input(shape, f, dy, strides, padding, oracle)
    assert au.rel_err(dx, oracle.conv2d_backprop_input(shape, f, dy, strides, padding)) < TOL_TF32
    dw = au.conv2d_backprop_filter(x, fshape, dy, strides, padding, oracle)
    assert au.rel_err(dw, oracle.conv2d_backprop_filter(x, fshape, dy, strides, padding)) < TOL_TF32


HALO_SHAPES = [
    # unit-stride shapes on the halo-tile kernel (conv_halo.cu): tile tails, several tiles per
    # image in both directions, VALID / SAME, 1 and 2 channel blocks, several N blocks, odd sizes
    ((3, 14, 14, 32), (5, 5, 32, 64), "SAME"),       # LeNet conv2: one image = one work item
    ((9, 14, 14, 64), (5, 5, 64, 32), "SAME"),       # its input gradient's shape (2 channel blocks)
    ((2, 40, 70, 32), (3, 3, 32, 32), "SAME"),       # several row bands per image
    ((1, 9, 300, 32), (3, 3, 32, 64), "VALID"),      # wider than one 256-pixel box: column tiles
    ((5, 11, 13, 32), (2, 4, 32, 96), "SAME"),       # even filter sizes: asymmetric SAME padding
    ((2, 8, 8, 64), (1, 3, 64, 320), "SAME"),        # K > 256: two N blocks, the second ragged
    ((7, 6, 6, 32), (6, 6, 32, 32), "VALID"),        # filter covers the image: 1x1 output
    ((37, 5, 5, 32), (3, 3, 32, 32), "SAME"),        # many small items: cluster tail padding
]


@pytest.mark.parametrize("shape,fshape,padding", HALO_SHAPES)
def test_conv_halo_forward_and_input_gradient_vs_oracle(oracle, rng, shape, fshape, padding):
    x = rng.rand(*shape).astype(np.float32) - 0.5
    f = (rng.rand(*fshape).astype(np.float32) - 0.5)
    y_ref = oracle.conv2d(x, f, (1, 1), padding)
    assert au.rel_err(au.conv2d(x, f, (1, 1), padding, oracle), y_ref) < TOL_TF32
    if shape[3] % 32 == 0 and fshape[3] % 32 == 0:
        dy = rng.rand(*y_ref.shape).astype(np.float32) - 0.5
        dx = au.conv2d_backprop_input(shape, f, dy, (1, 1), padding, oracle)
        assert au.rel_err(dx, oracle.conv2d_backprop_input(shape, f, dy, (1, 1), padding)) < TOL_TF32
        dw = au.conv2d_backprop_filter(x, fshape, dy, (1, 1), padding, oracle)
        assert au.rel_err(dw, oracle.conv2d_backprop_filter(x, fshape, dy, (1, 1), padding)) < TOL_TF32


def test_conv_halo_integer_data_is_exact(oracle, rng):
    # integers are exact in tf32 and their sums exact in fp32: any mis-addressed tap shows up
    x = rng.randint(-3, 4, (4, 14, 14, 32)).astype(np.float32)
    f = rng.randint(-3, 4, (5, 5, 32, 64)).astype(np.float32)
    np.testing.assert_array_equal(au.conv2d(x, f, (1, 1), "SAME", oracle),
                                  oracle.conv2d(x, f, (1, 1), "SAME"))
    np.testing.assert_array_equal(au.conv2d(x, f, (1, 1), "VALID", oracle),
                                  oracle.conv2d(x, f, (1, 1), "VALID"))
    dy = rng.randint(-3, 4, (4, 14, 14, 64)).astype(np.float32)
    np.testing.assert_array_equal(
        au.conv2d_backprop_input(x.shape, f, dy, (1, 1), "SAME", oracle),
        oracle.conv2d_backprop_input(x.shape, f, dy, (1, 1), "SAME"))
    # filter gradient: every tap group (horizontal, vertical, single) addressed exactly
    np.testing.assert_array_equal(
        au.conv2d_backprop_filter(x, f.shape, dy, (1, 1), "SAME", oracle),
        oracle.conv2d_backprop_filter(x, f.shape, dy, (1, 1), "SAME"))
    dyv = rng.randint(-3, 4, (4, 10, 10, 64)).astype(np.float32)
    np.testing.assert_array_equal(
        au.conv2d_backprop_filter(x, f.shape, dyv, (1, 1), "VALID", oracle),
        oracle.conv2d_backprop_filter(x, f.shape, dyv, (1, 1), "VALID"))
    for fs in [(3, 3, 32, 32), (2, 4, 32, 64), (1, 7, 32, 32), (6, 1, 32, 32)]:
        fz = rng.randint(-3, 4, fs).astype(np.float32)
        yz = oracle.conv2d(x, fz, (1, 1), "SAME")
        dz = rng.randint(-2, 3, yz.shape).astype(np.float32)
        np.testing.assert_array_equal(
            au.conv2d_backprop_filter(x, fs, dz, (1, 1), "SAME", oracle),
            oracle.conv2d_backprop_filter(x, fs, dz, (1, 1), "SAME"))


def test_conv_halo_bf16(oracle, rng):
    x = oracle.truncate_to_bf16(rng.rand(3, 20, 20, 64).astype(np.float32) - 0.5)
    f = oracle.truncate_to_bf16((rng.rand(3, 3, 64, 128).astype(np.float32) - 0.5) * 0.2)
    y_ref = oracle.conv2d(x, f, (1, 1), "SAME")
    assert au.rel_err(au.conv2d(x, f, (1, 1), "SAME", oracle, bf16=True), y_ref) < TOL
    dy = oracle.truncate_to_bf16(rng.rand(*y_ref.shape).astype(np.float32) - 0.5)
    dx = au.conv2d_backprop_input(x.shape, f, dy, (1, 1), "SAME", oracle, bf16=True)
    assert au.rel_err(dx, oracle.conv2d_backprop_input(x.shape, f, dy, (1, 1), "SAME")) < TOL
    dw = au.conv2d_backprop_filter(x, f.shape, dy, (1, 1), "SAME", oracle, bf16=True)
    assert au.rel_err(dw, oracle.conv2d_backprop_filter(x, f.shape, dy, (1, 1), "SAME")) < TOL


def test_conv2d_empty_batch(oracle):
    out = au.conv2d(np.zeros((0, 2, 3, 3), np.float32), gu.iota([1, 1, 3, 3]), [1, 1], "VALID",
                    oracle)
    assert out.shape == (0, 2, 3, 3)


def test_conv2d_implicit_matches_explicit(oracle, rng, monkeypatch):
    # same convolution through the TMA-im2col implicit GEMM and through the patch-matrix path
    x = rng.rand(6, 14, 14, 32).astype(np.float32)
    f = (rng.rand(5, 5, 32, 64).astype(np.float32) - 0.5)
    implicit = au.conv2d(x, f, (1, 1), "SAME", oracle)
    # (the switch is read once per process, so compare against the oracle for the other path)
    assert au.rel_err(implicit, oracle.conv2d(x, f, (1, 1), "SAME")) < TOL_TF32
    xi = rng.randint(-3, 4, (2, 9, 9, 32)).astype(np.float32)  # integer data: exact in tf32
    fi = rng.randint(-3, 4, (3, 3, 32, 8)).astype(np.float32)
    np.testing.assert_array_equal(au.conv2d(xi, fi, (1, 1), "SAME", oracle),
                                  oracle.conv2d(xi, fi, (1, 1), "SAME"))
    np.testing.assert_array_equal(au.conv2d(xi, fi, (2, 2), "VALID", oracle),
                                  oracle.conv2d(xi, fi, (2, 2), "VALID"))


def test_conv2d_bf16(oracle, rng):
    x = oracle.truncate_to_bf16(rng.rand(4, 14, 14, 32).astype(np.float32))
    f = oracle.truncate_to_bf16(rng.rand(3, 3, 32, 64).astype(np.float32) - 0.5)
    got = au.conv2d(x, f, (1, 1), "SAME", oracle, bf16=True)
    assert au.rel_err(got, oracle.conv2d(x, f, (1, 1), "SAME")) < TOL


@pytest.mark.parametrize("cin,cout", [(3, 64), (64, 64), (64, 128), (128, 128)])
def test_conv_family_bf16_vgg_layers(oracle, rng, cin, cout):
    # BASELINE config 5's layer type: 3x3 SAME stride 1 in bf16 (fp32 accumulate); C=3 runs the
    # patch-matrix path, the 64-multiples the implicit GEMM (forward, input and filter gradient)
    shape, fshape = (4, 16, 16, cin), (3, 3, cin, cout)
    x = oracle.truncate_to_bf16(rng.rand(*shape).astype(np.float32) - 0.5)
    f = oracle.truncate_to_bf16((rng.rand(*fshape).astype(np.float32) - 0.5) * 0.2)
    y_ref = oracle.conv2d(x, f, (1, 1), "SAME")
    assert au.rel_err(au.conv2d(x, f, (1, 1), "SAME", oracle, bf16=True), y_ref) < TOL
    dy = oracle.truncate_to_bf16(rng.rand(*y_ref.shape).astype(np.float32) - 0.5)
    dx = au.conv2d_backprop_input(shape, f, dy, (1, 1), "SAME", oracle, bf16=True)
    assert au.rel_err(dx, oracle.conv2d_backprop_input(shape, f, dy, (1, 1), "SAME")) < TOL
    dw = au.conv2d_backprop_filter(x, fshape, dy, (1, 1), "SAME", oracle, bf16=True)
    assert au.rel_err(dw, oracle.conv2d_backprop_filter(x, fshape, dy, (1, 1), "SAME")) < TOL


def test_conv2d_full_size_vs_oracle(oracle, rng):
    # BASELINE config 3 full size (batch 512 LeNet conv2, the shapes bench.py times): forward,
    # input gradient and filter gradient against the CPU oracle, element-wise and in norm
    shape, fshape = (512, 14, 14, 32), (5, 5, 32, 64)
    x = rng.rand(*shape).astype(np.float32) - 0.5
    f = (rng.rand(*fshape).astype(np.float32) - 0.5) * 0.1
    y = au.conv2d(x, f, (1, 1), "SAME", oracle)
    y_ref = oracle.conv2d(x, f, (1, 1), "SAME")
    assert au.rel_err(y, y_ref) < TOL_TF32
    dy = rng.rand(*y.shape).astype(np.float32) - 0.5
    dx = au.conv2d_backprop_input(shape, f, dy, (1, 1), "SAME", oracle)
    assert au.rel_err(dx, oracle.conv2d_backprop_input(shape, f, dy, (1, 1), "SAME")) < TOL_TF32
    dw = au.conv2d_backprop_filter(x, fshape, dy, (1, 1), "SAME", oracle)
    assert au.rel_err(dw, oracle.conv2d_backprop_filter(x, fshape, dy, (1, 1), "SAME")) < TOL_TF32
    # and the adjoint identities <dy, conv(x,w)> == <dX, x> == <dW, w> (size-independent property)
    lhs = float(np.sum(dy.astype(np.float64) * y))
    scale = float(np.sqrt(np.sum(dy.astype(np.float64) ** 2) * np.sum(y.astype(np.float64) ** 2)))
    assert abs(lhs - float(np.sum(dx.astype(np.float64) * x))) < 2e-3 * scale
    assert abs(lhs - float(np.sum(dw.astype(np.float64) * f))) < 2e-3 * scale


def test_conv_c1_forward_matches_the_generic_first_layer_kernel_bit_for_bit(oracle, rng, monkeypatch):
    # same taps in the same order with fp32 FMAs: the specialised kernel changes the schedule only
    shape, fshape = (7, 28, 28, 1), (5, 5, 1, 32)
    x = rng.rand(*shape).astype(np.float32) - 0.5
    f = (rng.rand(*fshape).astype(np.float32) - 0.5) * 0.2
    got = au.conv2d(x, f, (1, 1), "SAME", oracle)
    monkeypatch.setenv("B200TF_CONV_NO_C1", "1")
    # the knob is read once per process: the generic kernel runs in a child
    import os, subprocess, sys, tempfile
    with tempfile.TemporaryDirectory() as tmp:
        np.save(os.path.join(tmp, "x.npy"), x)
        np.save(os.path.join(tmp, "f.npy"), f)
        code = ("import numpy as np, sys; sys.path.insert(0, 'tests'); import abi_util as au;"
                "x = np.load(r'%s/x.npy'); f = np.load(r'%s/f.npy');"
                "np.save(r'%s/y.npy', au.conv2d(x, f, (1, 1), 'SAME', None))" % (tmp, tmp, tmp))
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        subprocess.run([sys.executable, "-c", code], check=True, cwd=root)
        np.testing.assert_array_equal(got, np.load(os.path.join(tmp, "y.npy")))


def test_conv1_full_size_vs_oracle(oracle, rng):
    # LeNet conv1 at batch 512 (C_in = 1: the direct first-layer kernels), forward + filter gradient
    shape, fshape = (512, 28, 28, 1), (5, 5, 1, 32)
    x = rng.rand(*shape).astype(np.float32) - 0.5
    f = (rng.rand(*fshape).astype(np.float32) - 0.5) * 0.2
    y_ref = oracle.conv2d(x, f, (1, 1), "SAME")
    assert au.rel_err(au.conv2d(x, f, (1, 1), "SAME", oracle), y_ref) < TOL_TF32
    dy = rng.rand(*y_ref.shape).astype(np.float32) - 0.5
    dw = au.conv2d_backprop_filter(x, fshape, dy, (1, 1), "SAME", oracle)
    assert au.rel_err(dw, oracle.conv2d_backprop_filter(x, fshape, dy, (1, 1), "SAME")) < TOL_TF32


# =============================================================================== fused pool backward
@pytest.mark.parametrize("shape,ksize,padding", [((8, 28, 28, 32), (2, 2), "VALID"),
                                                 ((4, 14, 14, 64), (2, 2), "SAME"),
                                                 ((3, 9, 7, 32), (2, 2), "SAME"),   # clipped last windows
                                                 ((2, 12, 9, 128), (3, 3), "VALID"),
                                                 ((5, 6, 6, 4), (2, 3), "VALID")])
def test_max_pool_grad_relu_bias_grad_fused(oracle, rng, shape, ksize, padding):
    # = MaxPoolGrad -> ReluGrad(features = the pool input) -> BiasAddGrad, three oracle calls
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    x[rng.rand(*shape) < 0.3] = 0.0               # exact zeros: relu mask is strict (f > 0)
    x[0, :2, :2, :] = 0.25                         # ties: the first maximum wins
    oh, ow, _, _ = au.pool_geometry(shape, ksize, ksize, padding)
    g = rng.uniform(-1, 1, (shape[0], oh, ow, shape[3])).astype(np.float32)
    got = au.max_pool_grad_relu_bias_grad(x, g, ksize, ksize, padding)
    assert got is not None
    dx = oracle.max_pool_grad(x, g, ksize, ksize, padding)
    dy = oracle.relu_grad(dx, x)
    np.testing.assert_array_equal(got[0], dy)
    ref_db = dy.reshape(-1, shape[3]).astype(np.float64).sum(0)
    assert au.rel_err(got[1], ref_db) < 1e-5
    again = au.max_pool_grad_relu_bias_grad(x, g, ksize, ksize, padding)
    np.testing.assert_array_equal(got[1], again[1])  # ordered reduction


def test_max_pool_grad_relu_bias_grad_fused_bf16_and_limits(oracle, rng):
    shape = (4, 8, 8, 64)
    x = oracle.truncate_to_bf16(rng.uniform(-1, 1, shape).astype(np.float32))
    g = oracle.truncate_to_bf16(rng.uniform(-1, 1, (4, 4, 4, 64)).astype(np.float32))
    got = au.max_pool_grad_relu_bias_grad(x, g, (2, 2), (2, 2), "VALID", bf16=True)
    dy = oracle.relu_grad(oracle.max_pool_grad(x, g, (2, 2), (2, 2), "VALID"), x)
    np.testing.assert_array_equal(got[0], dy)
    assert au.rel_err(got[1], dy.reshape(-1, 64).astype(np.float64).sum(0)) < TOL
    # overlapping windows / windows that leave cells uncovered / odd channel counts: not fused
    xf = rng.uniform(-1, 1, (2, 9, 9, 32)).astype(np.float32)
    assert au.max_pool_grad_relu_bias_grad(xf, np.zeros((2, 4, 4, 32), np.float32), (3, 3), (2, 2), "VALID") is None
    assert au.max_pool_grad_relu_bias_grad(xf, np.zeros((2, 4, 4, 32), np.float32), (2, 2), (2, 2), "VALID") is None
    xo = rng.uniform(-1, 1, (2, 8, 8, 12)).astype(np.float32)
    assert au.max_pool_grad_relu_bias_grad(xo, np.zeros((2, 4, 4, 12), np.float32), (2, 2), (2, 2), "VALID") is None


# =============================================================================== glue ops
def test_multi_tensor_sgd_is_bit_identical_to_per_variable_updates(rng):
    import ctypes
    L = au.lib()
    sizes = [1024 * 1024, 1024, 7, 333 * 1001]
    vars_ = [rng.randn(n).astype(np.float32) for n in sizes]
    deltas = [rng.randn(n).astype(np.float32) for n in sizes]
    alphas = [au.dev(np.array([a], np.float32)) for a in (0.01, 0.5, 1.0, 0.125)]
    one = [au.dev(v) for v in vars_]
    many = [au.dev(v) for v in vars_]
    dd = [au.dev(d) for d in deltas]
    for v, a, d, n in zip(one, alphas, dd, sizes):
        au.call(L.b200_apply_gradient_descent, 1, v.data_ptr(), a.data_ptr(), d.data_ptr(), n, au.stream())
    k = len(sizes)
    au.call(L.b200_apply_gradient_descent_multi, 1, k,
            (ctypes.c_void_p * k)(*[t.data_ptr() for t in many]),
            (ctypes.c_void_p * k)(*[t.data_ptr() for t in alphas]),
            (ctypes.c_void_p * k)(*[t.data_ptr() for t in dd]),
            (ctypes.c_int64 * k)(*sizes), au.stream())
    for a, b in zip(one, many):
        np.testing.assert_array_equal(au.host(a), au.host(b))


def test_apply_gradient_descent_add_n_scale_sum(oracle, rng):
    import torch
    L = au.lib()
    n = 1024 * 1024 + 5
    var = rng.randn(n).astype(np.float32)
    delta = rng.randn(n).astype(np.float32)
    dv, dd = au.dev(var), au.dev(delta)
    alpha = au.dev(np.array([0.01], np.float32))
    au.call(L.b200_apply_gradient_descent, 1, dv.data_ptr(), alpha.data_ptr(), dd.data_ptr(), n,
            au.stream())
    prod = au.empty((n,))
    au.call(L.b200_mul, 1, dd.data_ptr(), alpha.data_ptr(), prod.data_ptr(), n, 1, au.stream())
    np.testing.assert_array_equal(au.host(prod), delta * np.float32(0.01))
    au.call(L.b200_mul, 1, dd.data_ptr(), dd.data_ptr(), prod.data_ptr(), n, 0, au.stream())
    np.testing.assert_array_equal(au.host(prod), delta * delta)
    au.call(L.b200_add, 1, dd.data_ptr(), alpha.data_ptr(), prod.data_ptr(), n, 1, au.stream())
    np.testing.assert_array_equal(au.host(prod), delta + np.float32(0.01))
    au.call(L.b200_add, 1, dd.data_ptr(), dd.data_ptr(), prod.data_ptr(), n, 0, au.stream())
    np.testing.assert_array_equal(au.host(prod), delta + delta)
    np.testing.assert_allclose(au.host(dv), oracle.apply_gradient_descent(var, 0.01, delta),
                               rtol=1e-6, atol=1e-7)
    import ctypes
    ins = [au.dev(rng.randn(1000).astype(np.float32)) for _ in range(3)]
    ptrs = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in ins])
    out = au.empty((1000,))
    au.call(L.b200_add_n, 1, ptrs, 3, out.data_ptr(), 1000, au.stream())
    np.testing.assert_array_equal(au.host(out), (au.host(ins[0]) + au.host(ins[1])) + au.host(ins[2]))
    sc = au.empty((1000,))
    au.call(L.b200_scale, 1, ins[0].data_ptr(), 0.25, sc.data_ptr(), 1000, au.stream())
    np.testing.assert_array_equal(au.host(sc), au.host(ins[0]) * np.float32(0.25))
    tot = au.empty((1,))
    au.call(L.b200_reduce_sum, 1, dd.data_ptr(), 1.0 / n, tot.data_ptr(), n, au.stream())
    np.testing.assert_allclose(au.host(tot)[0], delta.astype(np.float64).mean(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("shape", [(3, 32, 49), (2, 1, 5), (4, 96, 1000), (1, 33, 31), (5, 7, 1)])
def test_batched_transpose_bit_exact(rng, shape):
    # pure data movement: the NCHW <-> NHWC layout change must be bit-exact (fp32 and bf16)
    L = au.lib()
    b, r, c = shape
    x = rng.randn(*shape).astype(np.float32)
    dx, out = au.dev(x), au.empty((b, c, r))
    au.call(L.b200_batched_transpose, 1, dx.data_ptr(), out.data_ptr(), b, r, c, au.stream())
    np.testing.assert_array_equal(au.host(out), np.swapaxes(x, 1, 2))
    xb = au.dev(x, True)
    outb = au.empty((b, c, r), au.tdt(True))
    au.call(L.b200_batched_transpose, 14, xb.data_ptr(), outb.data_ptr(), b, r, c, au.stream())
    np.testing.assert_array_equal(au.host(outb), np.swapaxes(au.host(xb), 1, 2))


def test_stream_event_memcpy_shim():
    import ctypes
    L = au.lib()
    s, e0, e1 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    assert L.b200_stream_create(ctypes.byref(s)) == 0
    assert L.b200_event_create(ctypes.byref(e0)) == 0 and L.b200_event_create(ctypes.byref(e1)) == 0
    d, h = ctypes.c_void_p(), ctypes.c_void_p()
    n = 1 << 20
    assert L.b200_malloc(ctypes.byref(d), n) == 0 and L.b200_host_malloc(ctypes.byref(h), n) == 0
    src = (ctypes.c_ubyte * n).from_address(h.value)
    for i in range(0, n, 4099):
        src[i] = i % 251
    assert L.b200_event_record(e0, s) == 0
    assert L.b200_memcpy_h2d_async(d, h, n, s) == 0
    assert L.b200_memset_async(h, 0, 0, s) == 0
    back = (ctypes.c_ubyte * n)()
    hb = ctypes.c_void_p()
    assert L.b200_host_malloc(ctypes.byref(hb), n) == 0
    assert L.b200_memcpy_d2h_async(hb, d, n, s) == 0
    assert L.b200_event_record(e1, s) == 0
    assert L.b200_stream_synchronize(s) == 0
    assert L.b200_event_query(e1) == 0
    ms = ctypes.c_float()
    assert L.b200_event_elapsed_ms(e0, e1, ctypes.byref(ms)) == 0 and ms.value >= 0
    got = (ctypes.c_ubyte * n).from_address(hb.value)
    assert all(got[i] == i % 251 for i in range(0, n, 4099))
    free_b, total_b = ctypes.c_size_t(), ctypes.c_size_t()
    assert L.b200_mem_info(ctypes.byref(free_b), ctypes.byref(total_b)) == 0 and total_b.value > 0
    for fn, p in ((L.b200_free, d), (L.b200_host_free, h), (L.b200_host_free, hb),
                  (L.b200_event_destroy, e0), (L.b200_event_destroy, e1), (L.b200_stream_destroy, s)):
        assert fn(p) == 0
    del back
