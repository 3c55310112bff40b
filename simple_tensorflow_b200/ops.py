"""Op constructors, symbolic gradients and SGD for the hot-path ops.

Mirrors the slices of the reference's Python front-end that build "forward + backward" graphs:
  python/ops/math_ops.py / nn_ops.py wrappers (matmul, bias_add, relu, conv2d, max_pool, ...)
  python/ops/math_grad.py:774-794 (_MatMulGrad), python/ops/nn_grad.py:180-204 (_BiasAddGrad),
  :267-269 (_ReluGrad), :323-333 (_SoftmaxCrossEntropyWithLogitsGrad), :363-374 (_Conv2DGrad),
  :430-438 (_MaxPoolGrad), python/ops/gradients_impl.py (reverse accumulation with AddN),
  python/training/gradient_descent.py (ApplyGradientDescent per variable).
Shapes are tracked statically here (the C layer checks them again at run time).
"""
import numpy as np

from . import client
from .client import (Graph, HostTensor, Operation, Output, Session, float32, int32, int64, bfloat16,
                     float16)

_default_graph = None


def get_default_graph():
    global _default_graph
    if _default_graph is None:
        _default_graph = Graph()
    return _default_graph


def reset_default_graph():
    global _default_graph
    _default_graph = Graph()
    return _default_graph


def _g(*outs):
    for o in outs:
        if isinstance(o, Output):
            return o.graph
    return get_default_graph()


def _shape(t):
    return t.graph.shapes.get(t.name)


def _set_shape(t, shape):
    t.graph.shapes[t.name] = tuple(int(s) for s in shape) if shape is not None else None
    return t


def _bf16_bits(a):
    """fp32 -> bfloat16 bit patterns by TRUNCATION (framework/bfloat16.cc:20-31)."""
    return (np.ascontiguousarray(a, np.float32).view(np.uint32) >> 16).astype(np.uint16)


# ------------------------------------------------------------------ sources
def placeholder(dtype, shape=None, name=None):
    g = get_default_graph()
    attrs = {"dtype": ("type", dtype)}
    if shape is not None:
        attrs["shape"] = ("shape", list(shape))
    op = g.create_op("Placeholder", [], attrs, name or "Placeholder")
    return _set_shape(op.outputs[0], shape)


def constant(value, dtype=None, shape=None, name=None):
    g = get_default_graph()
    arr = np.asarray(value)
    if dtype is None:
        dtype = {"f": float32, "i": int32}.get(arr.dtype.kind, float32)
        if arr.dtype == np.int64:
            dtype = int64
    if shape is not None:
        arr = np.broadcast_to(arr, shape)
    if dtype == bfloat16:
        host = HostTensor.allocate(bfloat16, arr.shape)
        host.numpy()[...] = _bf16_bits(arr)
    else:
        host = HostTensor.from_numpy(arr.astype(client._NP_OF[dtype]), dtype)
    op = g.create_op("Const", [], {"value": ("tensor", host), "dtype": ("type", dtype)},
                     name or "Const")
    return _set_shape(op.outputs[0], arr.shape)


class Variable:
    """python/ops/variables.py Variable: VariableV2 + Assign(initial_value)."""

    def __init__(self, initial_value, dtype=float32, name=None):
        g = get_default_graph()
        arr = np.asarray(initial_value, np.float32)
        base = name or "Variable"
        self.op = g.create_op("VariableV2", [], {"shape": ("shape", list(arr.shape)),
                                                 "dtype": ("type", dtype)}, base)
        self.ref = _set_shape(self.op.outputs[0], arr.shape)
        self.initial_value = constant(arr, dtype, name=self.op.name + "/initial_value")
        self.initializer = g.create_op("Assign", [self.ref, self.initial_value],
                                       {"T": ("type", dtype)}, self.op.name + "/Assign")
        self.dtype = dtype
        self.shape = tuple(arr.shape)
        g.variables.append(self)

    @classmethod
    def from_imported(cls, ref):
        """Wrap the output of an imported VariableV2 node (Graph.import_graph_def) so that
        optimizers can update it; no node is added.  Its initializer is whatever Assign the
        imported graph carries (e.g. the node named "<var>/Assign")."""
        if ref.op.type != "VariableV2":
            raise ValueError("%s is a %s, not a VariableV2" % (ref.op.name, ref.op.type))
        v = cls.__new__(cls)
        v.op, v.ref, v.dtype = ref.op, ref, ref.dtype
        v.shape = _shape(ref)
        v.initial_value = v.initializer = None
        return v

    def value(self):
        return self.ref

    def assign(self, value):
        return self.ref.graph.create_op("Assign", [self.ref, value], {"T": ("type", self.dtype)},
                                        self.op.name + "/Assign").outputs[0]


def _val(x):
    return x.ref if isinstance(x, Variable) else x


def global_variables_initializer():
    g = get_default_graph()
    return group(*[v.initializer for v in g.variables], name="init")


def group(*ops_, name=None):
    """control_flow_ops.group: a NoOp with control dependencies on every input."""
    g = get_default_graph()
    deps = [o.op if isinstance(o, Output) else o for o in ops_]
    return g.create_op("NoOp", [], {}, name or "group_deps", control_inputs=deps)


# ------------------------------------------------------------------ math / nn wrappers
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _val(a), _val(b)
    op = _g(a).create_op("MatMul", [a, b], {"T": ("type", a.dtype), "transpose_a": transpose_a,
                                            "transpose_b": transpose_b}, name or "MatMul")
    sa, sb = _shape(a), _shape(b)
    shape = None
    if sa and sb:
        shape = (sa[1] if transpose_a else sa[0], sb[0] if transpose_b else sb[1])
    return _set_shape(op.outputs[0], shape)


def batch_matmul(x, y, adj_x=False, adj_y=False, name=None):
    x, y = _val(x), _val(y)
    op = _g(x).create_op("BatchMatMul", [x, y], {"T": ("type", x.dtype), "adj_x": adj_x,
                                                 "adj_y": adj_y}, name or "BatchMatMul")
    sx, sy = _shape(x), _shape(y)
    shape = None
    if sx and sy:
        shape = tuple(sx[:-2]) + (sx[-1] if adj_x else sx[-2], sy[-2] if adj_y else sy[-1])
    return _set_shape(op.outputs[0], shape)


def bias_add(value, bias, data_format="NHWC", name=None):
    value, bias = _val(value), _val(bias)
    op = _g(value).create_op("BiasAdd", [value, bias],
                             {"T": ("type", value.dtype), "data_format": data_format},
                             name or "BiasAdd")
    return _set_shape(op.outputs[0], _shape(value))


def relu(features, name=None):
    features = _val(features)
    op = _g(features).create_op("Relu", [features], {"T": ("type", features.dtype)}, name or "Relu")
    return _set_shape(op.outputs[0], _shape(features))


def softmax(logits, name=None):
    op = _g(logits).create_op("Softmax", [logits], {"T": ("type", logits.dtype)}, name or "Softmax")
    return _set_shape(op.outputs[0], _shape(logits))


def log_softmax(logits, name=None):
    op = _g(logits).create_op("LogSoftmax", [logits], {"T": ("type", logits.dtype)},
                              name or "LogSoftmax")
    return _set_shape(op.outputs[0], _shape(logits))


def softmax_cross_entropy_with_logits(logits, labels, name=None):
    """-> per-example loss [batch] (nn_ops.py softmax_cross_entropy_with_logits, rank-2 case)."""
    op = _g(logits).create_op("SoftmaxCrossEntropyWithLogits", [logits, labels],
                              {"T": ("type", logits.dtype)}, name or "SoftmaxCrossEntropyWithLogits")
    s = _shape(logits)
    _set_shape(op.outputs[1], s)
    return _set_shape(op.outputs[0], (s[0],) if s else None)


def _windowed(in_size, filt, stride, padding):
    # framework/common_shape_fns.cc:19-47
    if padding == "VALID":
        return (in_size - filt + stride) // stride
    return (in_size + stride - 1) // stride


def _hw(data_format):
    """Positions of (H, W, C) in a 4-D activation of this data_format."""
    if data_format == "NCHW":
        return 2, 3, 1
    if data_format == "NHWC":
        return 1, 2, 3
    raise ValueError("data_format must be NHWC or NCHW")


def _spatial(shape, out_h, out_w, channels, data_format):
    return (shape[0], channels, out_h, out_w) if data_format == "NCHW" else (shape[0], out_h, out_w,
                                                                             channels)


def conv2d(input, filter, strides, padding, data_format="NHWC", name=None):  # noqa: A002
    input, filter = _val(input), _val(filter)
    h, w, _ = _hw(data_format)
    op = _g(input).create_op("Conv2D", [input, filter],
                             {"T": ("type", input.dtype), "strides": ("ints", list(strides)),
                              "padding": padding, "data_format": data_format}, name or "Conv2D")
    si, sf = _shape(input), _shape(filter)
    shape = None
    if si and sf:
        shape = _spatial(si, _windowed(si[h], sf[0], strides[h], padding),
                         _windowed(si[w], sf[1], strides[w], padding), sf[3], data_format)
    return _set_shape(op.outputs[0], shape)


def max_pool(value, ksize, strides, padding, data_format="NHWC", name=None):
    h, w, c = _hw(data_format)
    op = _g(value).create_op("MaxPool", [value],
                             {"T": ("type", value.dtype), "ksize": ("ints", list(ksize)),
                              "strides": ("ints", list(strides)), "padding": padding,
                              "data_format": data_format}, name or "MaxPool")
    s = _shape(value)
    shape = None
    if s:
        shape = _spatial(s, _windowed(s[h], ksize[h], strides[h], padding),
                         _windowed(s[w], ksize[w], strides[w], padding), s[c], data_format)
    return _set_shape(op.outputs[0], shape)


def reshape(tensor, shape, name=None):
    tensor = _val(tensor)
    shape_t = constant(np.asarray(shape, np.int32), int32)
    op = _g(tensor).create_op("Reshape", [tensor, shape_t], {"T": ("type", tensor.dtype)},
                              name or "Reshape")
    s = _shape(tensor)
    out = list(shape)
    if s is not None and -1 in out:
        known = int(np.prod([d for d in out if d != -1], dtype=np.int64))
        out[out.index(-1)] = int(np.prod(s, dtype=np.int64)) // max(known, 1)
    return _set_shape(op.outputs[0], out)


def cast(x, dtype, name=None):
    x = _val(x)
    op = _g(x).create_op("Cast", [x], {"SrcT": ("type", x.dtype), "DstT": ("type", dtype)},
                         name or "Cast")
    return _set_shape(op.outputs[0], _shape(x))


def argmax(input, axis, name=None):  # noqa: A002
    input = _val(input)
    dim = constant(np.asarray(axis, np.int32), int32)
    op = _g(input).create_op("ArgMax", [input, dim], {"T": ("type", input.dtype)}, name or "ArgMax")
    s = _shape(input)
    if s is not None:
        ax = axis % len(s)
        s = tuple(d for i, d in enumerate(s) if i != ax)
    return _set_shape(op.outputs[0], s)


def _reduce(op_type, x, axis, keep_dims, name):
    x = _val(x)
    s = _shape(x)
    rank = len(s)
    if axis is None:
        axes = list(range(rank))
    else:
        axes = [int(a) for a in (axis if isinstance(axis, (list, tuple)) else [axis])]
    idx = constant(np.asarray(axes, dtype=np.int32).reshape(len(axes)), int32)
    op = _g(x).create_op(op_type, [x, idx], {"T": ("type", x.dtype), "keep_dims": bool(keep_dims)},
                         name or op_type)
    red = {a % rank for a in axes} if rank else set()
    out = tuple((1 if d in red else s[d]) for d in range(rank) if keep_dims or d not in red)
    return _set_shape(op.outputs[0], out)


def reduce_mean(x, axis=None, keep_dims=False, name=None):
    """math_ops.reduce_mean (axis=None: over all axes -> scalar)."""
    return _reduce("Mean", x, axis, keep_dims, name)


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    """math_ops.reduce_sum (axis=None: over all axes -> scalar)."""
    return _reduce("Sum", x, axis, keep_dims, name)


def multiply(x, y, name=None):
    op = _g(x).create_op("Mul", [x, y], {"T": ("type", x.dtype)}, name or "Mul")
    sx, sy = _shape(x), _shape(y)
    return _set_shape(op.outputs[0], sx if sx and int(np.prod(sx)) != 1 else sy)


def add(x, y, name=None):
    """math_ops.add for equal shapes or a scalar operand."""
    x, y = _val(x), _val(y)
    op = _g(x).create_op("Add", [x, y], {"T": ("type", x.dtype)}, name or "Add")
    sx, sy = _shape(x), _shape(y)
    return _set_shape(op.outputs[0], sx if sx not in (None, ()) else sy)


def import_graph_def(graph_def, name=""):
    """importer.py import_graph_def into the default graph: serialized GraphDef bytes (e.g.
    written by real TensorFlow 1.0) -> {prefixed name: Operation}."""
    return get_default_graph().import_graph_def(graph_def, name)


def add_n(inputs, name=None):
    if len(inputs) == 1:
        return inputs[0]
    op = _g(inputs[0]).create_op("AddN", [list(inputs)],
                                 {"N": len(inputs), "T": ("type", inputs[0].dtype)},
                                 name or "AddN")
    return _set_shape(op.outputs[0], _shape(inputs[0]))


def identity(x, name=None):
    x = _val(x)
    op = _g(x).create_op("Identity", [x], {"T": ("type", x.dtype)}, name or "Identity")
    return _set_shape(op.outputs[0], _shape(x))


def all_reduce(var, scale=1.0, name=None):
    """Additive op: in-place NCCL sum of a variable-backed buffer across replicas."""
    ref = _val(var)
    op = _g(ref).create_op("B200AllReduce", [ref], {"T": ("type", ref.dtype), "scale": float(scale)},
                           name or "B200AllReduce")
    return _set_shape(op.outputs[0], _shape(ref))


def all_reduce_n(tensors, scale=1.0, name=None):
    """Additive op: ONE NCCL all-reduce (sum) over a packed arena of `tensors`, times `scale`."""
    tensors = list(tensors)
    op = _g(tensors[0]).create_op("B200AllReduceN", [tensors],
                                  {"N": len(tensors), "T": ("type", tensors[0].dtype),
                                   "scale": float(scale)}, name or "B200AllReduceN")
    return [_set_shape(o, _shape(t)) for o, t in zip(op.outputs, tensors)]


# ------------------------------------------------------------------ gradients
_GRAD = {}


def _register_gradient(op_type):
    def deco(fn):
        _GRAD[op_type] = fn
        return fn
    return deco


@_register_gradient("MatMul")
def _matmul_grad(op, grad):
    # math_grad.py:774-794
    a, b = op.inputs
    ta, tb = op.attrs["transpose_a"], op.attrs["transpose_b"]
    if not ta and not tb:
        return matmul(grad, b, transpose_b=True), matmul(a, grad, transpose_a=True)
    if not ta and tb:
        return matmul(grad, b), matmul(grad, a, transpose_a=True)
    if ta and not tb:
        return matmul(b, grad, transpose_b=True), matmul(a, grad)
    return (matmul(b, grad, transpose_a=True, transpose_b=True),
            matmul(grad, a, transpose_a=True, transpose_b=True))


@_register_gradient("BatchMatMul")
def _batch_matmul_grad(op, grad):
    # math_grad.py:871-894
    x, y = op.inputs
    adj_x, adj_y = op.attrs["adj_x"], op.attrs["adj_y"]
    if not adj_x:
        if not adj_y:
            return batch_matmul(grad, y, False, True), batch_matmul(x, grad, True, False)
        return batch_matmul(grad, y, False, False), batch_matmul(grad, x, True, False)
    if not adj_y:
        return batch_matmul(y, grad, False, True), batch_matmul(x, grad, False, False)
    return batch_matmul(y, grad, True, True), batch_matmul(grad, x, True, True)


@_register_gradient("Add")
def _add_grad(op, grad):
    # math_grad.py:592-600 for the case without broadcasting (equal static shapes)
    x, y = op.inputs
    if _shape(x) is None or _shape(x) != _shape(y):
        raise NotImplementedError("gradient of a broadcasting Add is outside the hot path")
    return grad, grad


@_register_gradient("Mul")
def _mul_grad(op, grad):
    # math_grad.py _MulGrad without broadcasting; a scalar operand only receives the other side's
    # gradient (its own would need a full reduction and is not asked for on the hot path)
    x, y = op.inputs
    sx, sy = _shape(x), _shape(y)
    if sx is not None and sx == sy:
        return multiply(grad, y), multiply(grad, x)
    if sy is not None and int(np.prod(sy)) == 1:
        return multiply(grad, y), None
    if sx is not None and int(np.prod(sx)) == 1:
        return None, multiply(grad, x)
    raise NotImplementedError("gradient of a broadcasting Mul is outside the hot path")


@_register_gradient("Cast")
def _cast_grad(op, grad):
    # math_grad.py:942-953: float types cast the gradient back
    return (cast(grad, op.inputs[0].dtype),)


@_register_gradient("BiasAdd")
def _bias_add_grad(op, grad):
    # nn_grad.py:180-204: (received_grad, BiasAddGrad(received_grad))
    fmt = op.attrs.get("data_format", "NHWC")
    g = op.graph.create_op("BiasAddGrad", [grad], {"T": ("type", grad.dtype), "data_format": fmt},
                           "BiasAddGrad")
    # GetBiasValueDims (bias_op.cc:140-150): NCHW puts the channel third from last
    cdim = -3 if fmt == "NCHW" and _shape(grad) and len(_shape(grad)) > 2 else -1
    return grad, _set_shape(g.outputs[0], (_shape(grad)[cdim],) if _shape(grad) else None)


@_register_gradient("Relu")
def _relu_grad(op, grad):
    # nn_grad.py:267-269: ReluGrad(grad, op.outputs[0])
    g = op.graph.create_op("ReluGrad", [grad, op.outputs[0]], {"T": ("type", grad.dtype)}, "ReluGrad")
    return (_set_shape(g.outputs[0], _shape(grad)),)


@_register_gradient("SoftmaxCrossEntropyWithLogits")
def _xent_grad(op, grad_loss, grad_backprop=None):
    # nn_grad.py:323-333: backprop * expand_dims(grad_loss, -1); labels get no gradient here.
    # The GPU Mul kernel broadcasts a one-element operand only, so a per-example grad_loss
    # ([batch] weights) is rejected while the graph is built instead of failing inside Run().
    sg = _shape(grad_loss)
    if sg is None or int(np.prod(sg, dtype=np.int64)) != 1:
        raise NotImplementedError(
            "gradient of SoftmaxCrossEntropyWithLogits with a per-example loss gradient %s needs a "
            "row-broadcasting Mul, which is outside the hot path; reduce the loss with reduce_mean"
            % (sg,))
    return multiply(op.outputs[1], grad_loss), None


def _full_reduction(op):
    x = op.inputs[0]
    out_elems = int(np.prod(_shape(op.outputs[0]) or (), dtype=np.int64))
    if out_elems != 1:
        raise NotImplementedError("gradient of a partial-axis %s is outside the hot path" % op.type)
    return x, int(np.prod(_shape(x), dtype=np.int64))


@_register_gradient("Mean")
def _mean_grad(op, grad):
    # math_grad.py _MeanGrad for a full reduction: grad / N tiled to the input shape.
    x, n = _full_reduction(op)
    scale = constant(np.float32(1.0 / n), x.dtype)
    g = multiply(grad, scale) if grad is not None else scale
    if x.op.type == "SoftmaxCrossEntropyWithLogits" and x.name == x.op.outputs[0].name:
        # the xent gradient multiplies its backprop by this scalar (one-element broadcast in the
        # Mul kernel; the executor folds it into the xent kernel): no [batch] tile is materialised
        return g, None
    # any other producer (Relu, MatMul, Reshape ...) needs a gradient of the input's own shape,
    # as _MeanGrad's tile() delivers: a constant 1/N tensor, scaled by the incoming scalar
    tile = constant(np.full(_shape(x), 1.0 / n, np.float32), x.dtype)
    return (multiply(tile, grad) if grad is not None else tile), None


@_register_gradient("Sum")
def _sum_grad(op, grad):
    # math_grad.py _SumGrad for a full reduction: grad tiled to the input shape
    x, _ = _full_reduction(op)
    if x.op.type == "SoftmaxCrossEntropyWithLogits" and x.name == x.op.outputs[0].name:
        return (grad if grad is not None else constant(np.float32(1.0), x.dtype)), None
    tile = constant(np.ones(_shape(x), np.float32), x.dtype)
    return (multiply(tile, grad) if grad is not None else tile), None


@_register_gradient("Conv2D")
def _conv2d_grad(op, grad):
    # nn_grad.py:363-374
    x, w = op.inputs
    attrs = {"T": ("type", x.dtype), "strides": op.attrs["strides"], "padding": op.attrs["padding"],
             "data_format": op.attrs.get("data_format", "NHWC")}
    g = op.graph
    in_sizes = constant(np.asarray(_shape(x), np.int32), int32)
    f_sizes = constant(np.asarray(_shape(w), np.int32), int32)
    dx = g.create_op("Conv2DBackpropInput", [in_sizes, w, grad], attrs, "Conv2DBackpropInput")
    dw = g.create_op("Conv2DBackpropFilter", [x, f_sizes, grad], attrs, "Conv2DBackpropFilter")
    return _set_shape(dx.outputs[0], _shape(x)), _set_shape(dw.outputs[0], _shape(w))


@_register_gradient("MaxPool")
def _max_pool_grad(op, grad):
    # nn_grad.py:430-438: MaxPoolGrad(op.inputs[0], op.outputs[0], grad)
    attrs = {"T": ("type", grad.dtype), "ksize": op.attrs["ksize"], "strides": op.attrs["strides"],
             "padding": op.attrs["padding"], "data_format": op.attrs.get("data_format", "NHWC")}
    g = op.graph.create_op("MaxPoolGrad", [op.inputs[0], op.outputs[0], grad], attrs, "MaxPoolGrad")
    return (_set_shape(g.outputs[0], _shape(op.inputs[0])),)


@_register_gradient("Reshape")
def _reshape_grad(op, grad):
    return reshape(grad, list(_shape(op.inputs[0]))), None


@_register_gradient("Identity")
def _identity_grad(op, grad):
    return (grad,)


def gradients(ys, xs):
    """gradients_impl.gradients: d(ys)/d(xs) by reverse accumulation.  ys: one scalar Output.
    xs: Outputs or Variables.  Only ops on a path from xs to ys get gradient nodes."""
    y = ys[0] if isinstance(ys, (list, tuple)) else ys
    targets = [_val(x) for x in xs]
    target_names = {t.name for t in targets}
    # forward reachability from xs
    reaches = {}

    def reach(t):
        if t.name in reaches:
            return reaches[t.name]
        r = t.name in target_names
        reaches[t.name] = r
        if not r:
            r = any([reach(i) for i in t.op.inputs])  # list: visit every input, no short circuit
            reaches[t.name] = r
        return r

    if not reach(y):
        return [None] * len(targets)
    # ops between xs and y in reverse topological order (creation order is topological)
    order = [op for op in y.graph.operations if any(reaches.get(o.name) for o in op.outputs)]
    pending = {y.name: [None]}  # None = implicit ones (scalar seed)
    result = {}
    for op in reversed(order):
        out_grads = []
        has_any = False
        for o in op.outputs:
            gl = pending.pop(o.name, [])
            if not gl:
                out_grads.append(None)
                continue
            has_any = True
            real = [g for g in gl if g is not None]
            out_grads.append(add_n(real) if real else None)
        if not has_any:
            continue
        for o, gsum in zip(op.outputs, out_grads):
            if o.name in target_names and (gsum is not None):
                result[o.name] = gsum
        if op.type in ("VariableV2", "Placeholder", "Const"):
            continue
        if op.type not in _GRAD:
            raise LookupError("No gradient defined for operation '%s' (op type: %s)" %
                              (op.name, op.type))
        seed_only = op.outputs[0].name == y.name and out_grads[0] is None
        if seed_only and op.type not in ("Mean", "Sum"):  # explicit ones seed (gradients_impl.py grad_ys=None)
            out_grads[0] = constant(np.ones(_shape(y) or (), np.float32), y.dtype)
            seed_only = False
        in_grads = _GRAD[op.type](op, *([None] if seed_only else out_grads[:1]))
        for inp, g in zip(op.inputs, in_grads):
            if g is not None and reaches.get(inp.name):
                pending.setdefault(inp.name, []).append(g)
    for name, gl in pending.items():
        if name in target_names:
            real = [g for g in gl if g is not None]
            if real:
                result[name] = add_n(real)
    return [result.get(t.name) for t in targets]


def _default_bucket_bytes():
    """Bucket size that matches how the session will run the exchange (see apply_gradients)."""
    import os
    if os.environ.get("B200TF_COLLECTIVE_OVERLAP") == "0":
        return None
    try:
        from . import _lib
        if _lib.load().b200_nvls_supported() == 1:
            return 4 << 20
    except Exception:
        pass
    return None


class GradientDescentOptimizer:
    """python/training/gradient_descent.py: one ApplyGradientDescent per variable."""

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)

    def compute_gradients(self, loss, var_list=None):
        var_list = var_list or get_default_graph().variables
        grads = gradients(loss, var_list)
        return list(zip(grads, var_list))

    def apply_gradients(self, grads_and_vars, name=None, num_replicas=1, bucket_bytes="auto"):
        g = get_default_graph()
        updates = []
        grads_and_vars = [(gr, v) for gr, v in grads_and_vars if gr is not None]
        if bucket_bytes == "auto":
            bucket_bytes = _default_bucket_bytes() if num_replicas > 1 else None
        if num_replicas > 1:
            # Replica data-parallel: gradients are averaged across replicas in buckets, filled
            # in the order backprop emits them (graph-construction order) and closed once they
            # hold `bucket_bytes`; each bucket is one collective that the executor starts as
            # soon as its gradients exist (under the rest of the backward pass when the session
            # runs collectives on their own stream, B200TF_COLLECTIVE_OVERLAP=1).
            # bucket_bytes=None: a single all-reduce of one contiguous gradient arena after the
            # whole backward pass.  "auto" (default): 4 MB buckets when the exchange will run as
            # the 16-CTA NVSwitch multicast kernel, which the session overlaps with the rest of the
            # backward pass inside the step's CUDA graph; one bucket otherwise (peer-IPC / NCCL
            # exchanges are faster exposed than overlapped: profiles/r01_notes.md, r02_notes.md).
            reduced = {}
            bucket, held = [], 0
            position = {id(op): i for i, op in enumerate(g.operations)}
            todo = sorted(grads_and_vars, key=lambda gv: position.get(id(gv[0].op), 0))
            for i, (gr, v) in enumerate(todo):
                if bucket and bucket[0][0].dtype != gr.dtype:
                    for (g0, v0), r in zip(bucket, all_reduce_n([b[0] for b in bucket], 1.0 / num_replicas)):
                        reduced[id(v0)] = r
                    bucket, held = [], 0
                bucket.append((gr, v))
                shape = _shape(gr)
                if shape is None:
                    shape = getattr(v, "shape", None) or ()
                held += int(np.prod(shape, dtype=np.int64)) * np.dtype(client._NP_OF[gr.dtype]).itemsize
                last = i == len(todo) - 1
                if last or (bucket_bytes is not None and held >= bucket_bytes):
                    for (g0, v0), r in zip(bucket, all_reduce_n([b[0] for b in bucket], 1.0 / num_replicas)):
                        reduced[id(v0)] = r
                    bucket, held = [], 0
            grads_and_vars = [(reduced[id(v)], v) for _, v in grads_and_vars]
        # what ApplyGradientDescent consumes (the replica average when num_replicas > 1):
        # fetchable, e.g. to check a data-parallel step against a full-batch reference
        self.applied_gradients = list(grads_and_vars)
        for grad, var in grads_and_vars:
            if grad is None:
                continue
            alpha = constant(np.float32(self.learning_rate), var.dtype)
            updates.append(g.create_op("ApplyGradientDescent", [var.ref, alpha, grad],
                                       {"T": ("type", var.dtype)},
                                       "GradientDescent/update_" + var.op.name))
        return group(*updates, name=name or "GradientDescent")

    def minimize(self, loss, var_list=None, name=None, num_replicas=1, bucket_bytes="auto"):
        return self.apply_gradients(self.compute_gradients(loss, var_list), name, num_replicas,
                                    bucket_bytes)
