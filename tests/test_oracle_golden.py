"""Pins the CPU oracle (oracle/oracle.c) against the reference's own known answers.

Sources (all under /root/reference/tensorflow/python/kernel_tests unless noted):
  conv_ops_test.py:298-406,487-567,637-705  -> tests/golden/conv_ops.json (extract_golden.py)
  pooling_ops_test.py:347-452,883-985       -> tests/golden/pooling_ops.json
  softmax_op_test.py:76-117, xent_op_test.py:96-131, relu_op_test.py:37-62,
  bias_op_test.py:48-134, matmul_op_test.py:48-85, argmax_op_test.py:28-68,
  cast_op_test.py:105-112, core/common_runtime/direct_session_test.cc:88-137
The NumPy expressions below are the same oracles those tests use (np.matrix product,
np.maximum, stable softmax, np.argmax ...).
"""
import numpy as np
import pytest

import golden_util as gu


# ------------------------------------------------------------------ padding arithmetic
@pytest.mark.parametrize("inp,filt,stride,padding,expect", [
    (3, 2, 1, "VALID", (2, 0, 0)), (3, 2, 2, "VALID", (1, 0, 0)), (7, 2, 3, "VALID", (2, 0, 0)),
    (3, 2, 2, "SAME", (2, 0, 1)), (4, 2, 3, "SAME", (2, 0, 1)), (28, 5, 1, "SAME", (28, 2, 2)),
    (8, 3, 2, "SAME", (4, 0, 1)), (5, 3, 1, "SAME", (5, 1, 1)), (224, 3, 1, "SAME", (224, 1, 1)),
    (2, 5, 1, "SAME", (2, 2, 2)), (1, 1, 2, "SAME", (1, 0, 0)),
])
def test_windowed_output_size(oracle, inp, filt, stride, padding, expect):
    # framework/common_shape_fns.cc:19-47: VALID out=(in-f+s)/s; SAME out=ceil(in/s),
    # pad_before = total/2, pad_after = the rest
    assert oracle.windowed_output_size(inp, filt, stride, padding) == expect


# ------------------------------------------------------------------ conv goldens
@pytest.mark.parametrize("case", gu.conv_cases("conv2d"), ids=gu.case_id)
def test_conv2d_golden(oracle, case):
    x = gu.iota(case["tensor_in_sizes"])
    f = gu.iota(case["filter_in_sizes"])
    out = oracle.conv2d(x, f, case["strides"], case["padding"])
    np.testing.assert_allclose(out.ravel(), np.asarray(case["expected"], np.float32),
                               rtol=1e-5, atol=1e-5)  # tol of conv_ops_test.py:292-295


@pytest.mark.parametrize("case", gu.conv_cases("conv2d_backprop_input"), ids=gu.case_id)
def test_conv2d_backprop_input_golden(oracle, case):
    f = gu.iota(case["filter_sizes"])
    dy = gu.iota(case["output_sizes"])
    out = oracle.conv2d_backprop_input(case["input_sizes"], f, dy, case["strides"],
                                       case["padding"])
    np.testing.assert_allclose(out.ravel(), np.asarray(case["expected"], np.float32),
                               rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("case", gu.conv_cases("conv2d_backprop_filter"), ids=gu.case_id)
def test_conv2d_backprop_filter_golden(oracle, case):
    x = gu.iota(case["input_sizes"])
    dy = gu.iota(case["output_sizes"])
    out = oracle.conv2d_backprop_filter(x, case["filter_sizes"], dy, case["strides"],
                                        case["padding"])
    np.testing.assert_allclose(out.ravel(), np.asarray(case["expected"], np.float32),
                               rtol=1e-5, atol=1e-4)


def _naive_conv(x, f, strides, padding, oracle):
    """6-loop reference of eigen_spatial_convolutions_test.cc:53-71, with TF's SAME/VALID."""
    n, h, w, c = x.shape
    r, s, _, k = f.shape
    oh, pt, _ = oracle.windowed_output_size(h, r, strides[0], padding)
    ow, pl, _ = oracle.windowed_output_size(w, s, strides[1], padding)
    out = np.zeros((n, oh, ow, k), np.float64)
    for i in range(oh):
        for j in range(ow):
            for a in range(r):
                for b in range(s):
                    ih, iw = i * strides[0] - pt + a, j * strides[1] - pl + b
                    if 0 <= ih < h and 0 <= iw < w:
                        out[:, i, j, :] += x[:, ih, iw, :].astype(np.float64) @ f[a, b].astype(np.float64)
    return out


@pytest.mark.parametrize("shape,fshape,strides,padding", [
    ((2, 9, 8, 3), (3, 3, 3, 5), (1, 1), "SAME"), ((2, 9, 8, 3), (3, 2, 3, 4), (2, 1), "VALID"),
    ((1, 7, 7, 2), (5, 5, 2, 3), (2, 2), "SAME"), ((3, 5, 6, 4), (1, 1, 4, 6), (1, 1), "VALID"),
])
def test_conv_family_consistency(oracle, rng, shape, fshape, strides, padding):
    x = rng.rand(*shape).astype(np.float32)
    f = rng.rand(*fshape).astype(np.float32)
    y = oracle.conv2d(x, f, strides, padding)
    np.testing.assert_allclose(y, _naive_conv(x, f, strides, padding, oracle), rtol=1e-5, atol=1e-5)
    # adjoint identities: <dy, conv(x, f)> == <dX, x> == <dW, f>
    dy = rng.rand(*y.shape).astype(np.float32)
    dx = oracle.conv2d_backprop_input(shape, f, dy, strides, padding)
    dw = oracle.conv2d_backprop_filter(x, fshape, dy, strides, padding)
    lhs = float(np.sum(dy.astype(np.float64) * y))
    assert abs(lhs - float(np.sum(dx.astype(np.float64) * x))) < 1e-4 * abs(lhs)
    assert abs(lhs - float(np.sum(dw.astype(np.float64) * f))) < 1e-4 * abs(lhs)


def test_conv2d_empty_batch(oracle):
    # conv_ops_test.py:310-317 testConv2DEmpty
    out = oracle.conv2d(np.zeros((0, 2, 3, 3), np.float32), gu.iota([1, 1, 3, 3]), [1, 1], "VALID")
    assert out.shape == (0, 2, 3, 3)


# ------------------------------------------------------------------ pooling goldens
@pytest.mark.parametrize("case", gu.pool_cases("max_pool"), ids=gu.case_id)
def test_max_pool_golden(oracle, case):
    x = gu.iota(case["input_sizes"])
    out = oracle.max_pool(x, case["ksize"][1:3], case["strides"][1:3], case["padding"])
    np.testing.assert_array_equal(out.ravel(), np.asarray(case["expected"], np.float32))


@pytest.mark.parametrize("case", gu.pool_cases("max_pool_grad"), ids=gu.case_id)
def test_max_pool_grad_direct_golden(oracle, case):
    # pooling_ops_test.py:883-985: ties go to the first maximum of the window
    x = np.asarray(case["input_data"], np.float32).reshape(case["input_sizes"])
    g = np.asarray(case["output_backprop"], np.float32).reshape(case["output_sizes"])
    out = oracle.max_pool_grad(x, g, [case["window_rows"], case["window_cols"]],
                               [case["row_stride"], case["col_stride"]], case["padding"])
    np.testing.assert_allclose(out.ravel(), np.asarray(case["expected_input_backprop"], np.float32),
                               rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ matmul
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (3, 5, 1), (5, 3, 5), (1, 5, 3), (128, 128, 128)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_matmul_numpy(oracle, rng, m, n, k, ta, tb):
    # matmul_op_test.py:48-85,224-231: N(-5, 5)-ish reals, np.matrix(a) * np.matrix(b)
    a = rng.normal(-5, 5, (k, m) if ta else (m, k)).astype(np.float32)
    b = rng.normal(-5, 5, (n, k) if tb else (k, n)).astype(np.float32)
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    out = oracle.matmul(a, b, ta, tb)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("m,n,k", [(67, 257, 5), (133, 300, 600), (6, 32, 256), (7, 33, 257),
                                   (200, 1024, 513)])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_matmul_blocking_edges_are_sequential_fp32_sums(oracle, rng, m, n, k, ta, tb):
    # the register-blocked GEMM (6 x 32 micro tiles, 66 x 256 x 256 blocks) must equal the plain
    # "c += a * b for k ascending" fp32 loop bit for bit at every tile / block edge
    a = rng.randn(*((k, m) if ta else (m, k))).astype(np.float32)
    b = rng.randn(*((n, k) if tb else (k, n))).astype(np.float32)
    A = a.T if ta else a
    B = b.T if tb else b
    ref = np.zeros((m, n), np.float32)
    for kk in range(k):   # separate multiply and add, ascending k: what -ffp-contract=off compiles to
        ref += (A[:, kk:kk + 1] * B[kk:kk + 1, :]).astype(np.float32)
    np.testing.assert_array_equal(oracle.matmul(a, b, ta, tb), ref)


def test_matmul_direct_session_known_answers(oracle):
    # core/common_runtime/direct_session_test.cc:54-108: a=[[3,2],[-1,0]], x=[[1],[1]] -> y=a*x
    a = np.array([[3, 2], [-1, 0]], np.float32)
    x = np.array([[1], [1]], np.float32)
    y = oracle.matmul(a, x)
    np.testing.assert_array_equal(y.ravel(), [5.0, -1.0])  # :107 expects y(0,0)=5.0, y(1,0)=-1.0


def test_matmul_zero_sizes(oracle):
    # matmul_op.cc:240-253: empty output -> nothing; k == 0 -> zeros
    assert oracle.matmul(np.zeros((0, 4), np.float32), np.zeros((4, 3), np.float32)).shape == (0, 3)
    out = oracle.matmul(np.zeros((2, 0), np.float32), np.zeros((0, 3), np.float32))
    np.testing.assert_array_equal(out, np.zeros((2, 3), np.float32))


@pytest.mark.parametrize("adj_x", [False, True])
@pytest.mark.parametrize("adj_y", [False, True])
def test_batch_matmul_numpy(oracle, rng, adj_x, adj_y):
    # batch_matmul_op_test.py: np.matmul with adj flags
    x = rng.randn(*((3, 7, 5) if adj_x else (3, 5, 7))).astype(np.float32)
    y = rng.randn(*((3, 4, 7) if adj_y else (3, 7, 4))).astype(np.float32)
    ref = np.matmul((x.transpose(0, 2, 1) if adj_x else x).astype(np.float64),
                    (y.transpose(0, 2, 1) if adj_y else y).astype(np.float64))
    np.testing.assert_allclose(oracle.batch_matmul(x, y, adj_x, adj_y), ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ element-wise / reductions
def test_bias_add_and_grad(oracle, rng):
    # bias_op_test.py:48-134: numpy broadcast add over the last dimension
    x = rng.rand(4, 3, 2, 5).astype(np.float32)
    b = rng.rand(5).astype(np.float32)
    np.testing.assert_array_equal(oracle.bias_add(x, b), x + b)
    np.testing.assert_allclose(oracle.bias_add_grad(x), x.reshape(-1, 5).sum(0), rtol=1e-6)


def test_relu_known_values(oracle):
    # relu_op_test.py:37-62: np.maximum(x, 0) on +-{0.1 .. 0.9}
    x = np.array([[-0.9, 0.7, -0.5, 0.3, -0.1], [0.1, -0.3, 0.5, -0.7, 0.9]], np.float32)
    np.testing.assert_array_equal(oracle.relu(x), np.maximum(x, 0))
    g = np.arange(1, 11, dtype=np.float32).reshape(2, 5)
    np.testing.assert_array_equal(oracle.relu_grad(g, x), g * (x > 0))
    # zero activation passes no gradient (relu_op_functor.h:51-56)
    np.testing.assert_array_equal(oracle.relu_grad(np.ones(3, np.float32),
                                                   np.array([0.0, -0.0, 1e-30], np.float32)),
                                  [0, 0, 1])


def test_softmax_known_answers(oracle):
    # softmax_op_test.py:76-95
    f = np.array([[1., 1., 1., 1.], [1., 2., 3., 4.]], np.float32)
    np.testing.assert_allclose(oracle.softmax(f), [[0.25, 0.25, 0.25, 0.25],
                                                   [0.0320586, 0.08714432, 0.23688282, 0.64391426]],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.softmax(f, log=True),
                               [[-1.386294] * 4, [-3.4401897, -2.4401897, -1.4401897, -0.4401897]],
                               rtol=1e-5, atol=1e-5)


def test_log_softmax_overflow(oracle):
    # softmax_op_test.py:102-117: [max, 1, 2, 3] -> [0, -max, -max, -max]
    mx = np.finfo(np.float32).max
    f = np.array([[1., 1., 1., 1.], [mx, 1., 2., 3.]], np.float32)
    np.testing.assert_allclose(oracle.softmax(f, log=True),
                               [[-1.386294] * 4, [0, -mx, -mx, -mx]], rtol=1e-5, atol=1e-5)


def test_xent_known_answers(oracle):
    # xent_op_test.py:96-131
    f = np.array([[1., 1., 1., 1.], [1., 2., 3., 4.]], np.float32)
    l = np.array([[0., 0., 0., 1.], [0., .5, .5, 0.]], np.float32)
    loss, bp = oracle.softmax_xent(f, l)
    np.testing.assert_allclose(bp, [[0.25, 0.25, 0.25, -0.75], [0.0321, -0.4129, -0.2632, 0.6439]],
                               rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(loss, [1.3862, 1.9401], rtol=1e-3, atol=1e-3)


def test_argmax_every_axis(oracle, rng):
    # argmax_op_test.py:28-68: [3,2,4,5,6] tensor, every axis -5..4, np.argmax (first on ties)
    x = rng.randn(3, 2, 4, 5, 6).astype(np.float32)
    for axis in range(-5, 5):
        np.testing.assert_array_equal(oracle.argmax(x, axis), np.argmax(x, axis=axis))
    xi = rng.randint(0, 3, (4, 7, 3)).astype(np.int32)  # many ties
    for axis in range(3):
        np.testing.assert_array_equal(oracle.argmax(xi, axis), np.argmax(xi, axis=axis))


def test_cast_bfloat16_is_truncation(oracle, rng):
    # framework/bfloat16.cc:20-50: keep the upper 16 bits; cast_op_test.py:105-112 round trip
    x = rng.randn(1000).astype(np.float32)
    b = oracle.cast_f32_to_bf16(x)
    np.testing.assert_array_equal(b, (x.view(np.uint32) >> 16).astype(np.uint16))
    back = oracle.cast_bf16_to_f32(b)
    np.testing.assert_array_equal(back.view(np.uint32), (x.view(np.uint32) >> 16) << 16)
    np.testing.assert_allclose(back, x, rtol=1 / 128.)
    # framework/bfloat16_test.cc:31-45 (Bfloat16Test.Conversion): a[i] = i + 1.25, |c - a| / a <= 1/128
    a = (np.arange(100) + 1.25).astype(np.float32)
    c = oracle.cast_bf16_to_f32(oracle.cast_f32_to_bf16(a))
    assert np.all(np.abs(c - a) / a <= 1.0 / 128)
    # just below the next bf16 value (1 + 2^-7): round-to-nearest would go up, truncation stays
    v = np.array([1.0 + 2 ** -7 - 2 ** -20], np.float32)
    assert oracle.cast_bf16_to_f32(oracle.cast_f32_to_bf16(v))[0] == np.float32(1.0)


def test_cast_numeric(oracle):
    x = np.array([-2.7, -0.5, 0.0, 0.5, 2.7, 1e6], np.float32)
    np.testing.assert_array_equal(oracle.cast(x, np.int32), x.astype(np.int32))
    np.testing.assert_array_equal(oracle.cast(x, np.int64), x.astype(np.int64))
    i = np.array([-(2 ** 31), -1, 0, 1, 2 ** 31 - 1], np.int32)
    np.testing.assert_array_equal(oracle.cast(i, np.float32), i.astype(np.float32))
    np.testing.assert_array_equal(oracle.cast(i, np.int64), i.astype(np.int64))
    j = np.array([-(2 ** 40), 5, 2 ** 33 + 1], np.int64)
    np.testing.assert_array_equal(oracle.cast(j, np.float32), j.astype(np.float32))
    np.testing.assert_array_equal(oracle.cast(j, np.int32), j.astype(np.int32))


def test_apply_gradient_descent(oracle):
    # training_ops.cc:410-412
    var = np.array([1.0, 2.0, 3.0], np.float32)
    np.testing.assert_allclose(oracle.apply_gradient_descent(var, 0.5, np.array([2.0, 2.0, -2.0])),
                               [0.0, 1.0, 4.0])
