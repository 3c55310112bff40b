"""Graph / Operation / Session on top of the C API (libb200tf_framework.so).

The role of the reference's tensorflow/python/framework/ops.py (Graph, Operation, Tensor) and
tensorflow/python/client/session.py (BaseSession.run :660-1012) -- reduced to what driving the
hot path needs, and talking to the same C entry points (TF_NewOperation / TF_FinishOperation /
TF_SessionRun, tensorflow/c/c_api.h) the reference reaches through SWIG
(python/client/tf_session_helper.cc:462,575).
"""
import ctypes
import weakref
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
FRAMEWORK_PATH = os.path.join(_HERE, "lib", "libb200tf_framework.so")

# TF_DataType <-> numpy.  bfloat16 has no numpy dtype: it travels as uint16 bit patterns.
TF_FLOAT, TF_INT32, TF_INT64, TF_BFLOAT16, TF_HALF = 1, 3, 9, 14, 19
_NP_OF = {TF_FLOAT: np.float32, TF_INT32: np.int32, TF_INT64: np.int64, TF_BFLOAT16: np.uint16,
          TF_HALF: np.float16}
_TF_OF = {np.dtype(np.float32): TF_FLOAT, np.dtype(np.int32): TF_INT32,
          np.dtype(np.int64): TF_INT64, np.dtype(np.float16): TF_HALF}
float32, int32, int64, bfloat16, float16 = TF_FLOAT, TF_INT32, TF_INT64, TF_BFLOAT16, TF_HALF

c_void_p, c_int, c_int64, c_size_t, c_char_p = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                 ctypes.c_size_t, ctypes.c_char_p)


class TF_Output(ctypes.Structure):
    _fields_ = [("oper", c_void_p), ("index", c_int)]


class TF_Input(ctypes.Structure):
    _fields_ = [("oper", c_void_p), ("index", c_int)]


class TF_AttrMetadata(ctypes.Structure):
    _fields_ = [("is_list", ctypes.c_ubyte), ("list_size", c_int64), ("type", c_int),
                ("total_size", c_int64)]


class TF_Buffer(ctypes.Structure):
    _fields_ = [("data", c_void_p), ("length", c_size_t), ("data_deallocator", c_void_p)]


class RunStats(ctypes.Structure):
    _fields_ = [("nodes_executed", c_int64), ("kernels_launched", c_int64),
                ("h2d_bytes", c_int64), ("d2h_bytes", c_int64),
                ("host_enqueue_us", c_int64), ("host_total_us", c_int64)]


_SIGS = {
    "TF_Version": (c_char_p, []),
    "TF_NewStatus": (c_void_p, []), "TF_DeleteStatus": (None, [c_void_p]),
    "TF_GetCode": (c_int, [c_void_p]), "TF_Message": (c_char_p, [c_void_p]),
    "TF_AllocateTensor": (c_void_p, [c_int, ctypes.POINTER(c_int64), c_int, c_size_t]),
    "TF_DeleteTensor": (None, [c_void_p]), "TF_TensorType": (c_int, [c_void_p]),
    "TF_NumDims": (c_int, [c_void_p]), "TF_Dim": (c_int64, [c_void_p, c_int]),
    "TF_TensorByteSize": (c_size_t, [c_void_p]), "TF_TensorData": (c_void_p, [c_void_p]),
    "TF_NewSessionOptions": (c_void_p, []), "TF_DeleteSessionOptions": (None, [c_void_p]),
    "TF_SetTarget": (None, [c_void_p, c_char_p]),
    "B200TF_SetGpuDevice": (None, [c_void_p, c_int]),
    "B200TF_SetGpuMemoryLimit": (None, [c_void_p, c_size_t]),
    "B200TF_SetCollective": (None, [c_void_p, c_void_p, c_int]),
    "TF_NewGraph": (c_void_p, []), "TF_DeleteGraph": (None, [c_void_p]),
    "TF_NewOperation": (c_void_p, [c_void_p, c_char_p, c_char_p]),
    "TF_SetDevice": (None, [c_void_p, c_char_p]),
    "TF_AddInput": (None, [c_void_p, TF_Output]),
    "TF_AddInputList": (None, [c_void_p, ctypes.POINTER(TF_Output), c_int]),
    "TF_AddControlInput": (None, [c_void_p, c_void_p]),
    "TF_SetAttrString": (None, [c_void_p, c_char_p, c_char_p, c_size_t]),
    "TF_SetAttrInt": (None, [c_void_p, c_char_p, c_int64]),
    "TF_SetAttrIntList": (None, [c_void_p, c_char_p, ctypes.POINTER(c_int64), c_int]),
    "TF_SetAttrFloat": (None, [c_void_p, c_char_p, ctypes.c_float]),
    "TF_SetAttrBool": (None, [c_void_p, c_char_p, ctypes.c_ubyte]),
    "TF_SetAttrType": (None, [c_void_p, c_char_p, c_int]),
    "TF_SetAttrShape": (None, [c_void_p, c_char_p, ctypes.POINTER(c_int64), c_int]),
    "TF_SetAttrTensor": (None, [c_void_p, c_char_p, c_void_p, c_void_p]),
    "TF_FinishOperation": (c_void_p, [c_void_p, c_void_p]),
    "TF_OperationName": (c_char_p, [c_void_p]), "TF_OperationOpType": (c_char_p, [c_void_p]),
    "TF_OperationNumOutputs": (c_int, [c_void_p]),
    "TF_OperationOutputType": (c_int, [TF_Output]),
    "TF_OperationNumInputs": (c_int, [c_void_p]),
    "TF_GraphOperationByName": (c_void_p, [c_void_p, c_char_p]),
    "TF_GraphNextOperation": (c_void_p, [c_void_p, ctypes.POINTER(c_size_t)]),
    "TF_OperationInput": (TF_Output, [TF_Input]),
    "TF_OperationNumControlInputs": (c_int, [c_void_p]),
    "TF_OperationGetControlInputs": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int]),
    "TF_OperationGetAttrMetadata": (TF_AttrMetadata, [c_void_p, c_char_p, c_void_p]),
    "TF_OperationGetAttrString": (None, [c_void_p, c_char_p, c_void_p, c_size_t, c_void_p]),
    "TF_OperationGetAttrInt": (None, [c_void_p, c_char_p, ctypes.POINTER(c_int64), c_void_p]),
    "TF_OperationGetAttrIntList": (None, [c_void_p, c_char_p, ctypes.POINTER(c_int64), c_int, c_void_p]),
    "TF_OperationGetAttrFloat": (None, [c_void_p, c_char_p, ctypes.POINTER(ctypes.c_float), c_void_p]),
    "TF_OperationGetAttrBool": (None, [c_void_p, c_char_p, ctypes.POINTER(ctypes.c_ubyte), c_void_p]),
    "TF_OperationGetAttrType": (None, [c_void_p, c_char_p, ctypes.POINTER(c_int), c_void_p]),
    "TF_OperationGetAttrShape": (None, [c_void_p, c_char_p, ctypes.POINTER(c_int64), c_int, c_void_p]),
    "TF_OperationGetAttrTensor": (None, [c_void_p, c_char_p, ctypes.POINTER(c_void_p), c_void_p]),
    "B200TF_OperationAttrNames": (c_void_p, [c_void_p]),
    "TF_NewBufferFromString": (c_void_p, [c_char_p, c_size_t]),
    "TF_NewBuffer": (c_void_p, []), "TF_DeleteBuffer": (None, [c_void_p]),
    "TF_GraphToGraphDef": (None, [c_void_p, c_void_p, c_void_p]),
    "TF_NewImportGraphDefOptions": (c_void_p, []),
    "TF_DeleteImportGraphDefOptions": (None, [c_void_p]),
    "TF_ImportGraphDefOptionsSetPrefix": (None, [c_void_p, c_char_p]),
    "TF_GraphImportGraphDef": (None, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "B200TF_GraphDefToText": (c_void_p, [c_char_p, c_size_t, c_void_p]),
    "TF_NewSession": (c_void_p, [c_void_p, c_void_p, c_void_p]),
    "TF_CloseSession": (None, [c_void_p, c_void_p]),
    "TF_DeleteSession": (None, [c_void_p, c_void_p]),
    "TF_SessionRun": (None, [c_void_p, c_void_p, ctypes.POINTER(TF_Output),
                             ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(TF_Output),
                             ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(c_void_p), c_int,
                             c_void_p, c_void_p]),
    "B200TF_SessionLastRunStats": (None, [c_void_p, ctypes.POINTER(RunStats)]),
    "B200TF_SessionStream": (c_void_p, [c_void_p]),
    "B200TF_SessionStageTensor": (c_void_p, [c_void_p, c_void_p, c_void_p]),
    "B200TF_ListRegisteredOps": (c_void_p, []),
    "B200TF_ListRegisteredKernels": (c_void_p, []),
    "TF_LoadLibrary": (c_void_p, [c_char_p, c_void_p]),
    "TF_DeleteLibraryHandle": (None, [c_void_p]),
}

_fw = None


def framework():
    """dlopen libb200tf_framework.so (after libb200tf.so) and attach prototypes."""
    global _fw
    if _fw is None:
        _lib.load()  # the kernel library first: the framework links against it
        if not os.path.exists(FRAMEWORK_PATH):
            raise ImportError(FRAMEWORK_PATH + " is missing: run __graft_entry__.build()")
        fw = ctypes.CDLL(FRAMEWORK_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(fw, name)
            fn.restype = res
            fn.argtypes = args
        _fw = fw
    return _fw


class OpError(RuntimeError):
    """A non-OK tensorflow::Status (python/framework/errors_impl.py OpError)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.error_code = code
        self.message = message


class _Status:
    def __init__(self):
        self.fw = framework()
        self.ptr = self.fw.TF_NewStatus()

    def check(self):
        code = self.fw.TF_GetCode(self.ptr)
        if code != 0:
            raise OpError(code, self.fw.TF_Message(self.ptr).decode("utf-8", "replace"))

    def __del__(self):
        try:
            self.fw.TF_DeleteStatus(self.ptr)
        except Exception:
            pass


def _free_cstr_list(ptr):
    s = ctypes.cast(ptr, c_char_p).value.decode()
    ctypes.CDLL(None).free(c_void_p(ptr))
    return [l for l in s.split("\n") if l]


def registered_ops():
    return _free_cstr_list(framework().B200TF_ListRegisteredOps())


def registered_kernels():
    return _free_cstr_list(framework().B200TF_ListRegisteredKernels())


# ------------------------------------------------------------------ tensors
class HostTensor:
    """Owns a TF_Tensor (pinned host memory when a GPU is present) and views it as numpy."""

    def __init__(self, ptr, owned=True):
        self.fw = framework()
        self.ptr = ptr
        self.owned = owned

    @classmethod
    def allocate(cls, dtype, shape):
        fw = framework()
        dims = (c_int64 * max(len(shape), 1))(*shape)
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        ptr = fw.TF_AllocateTensor(dtype, dims, len(shape), n * np.dtype(_NP_OF[dtype]).itemsize)
        if not ptr:
            raise MemoryError("TF_AllocateTensor failed")
        return cls(ptr)

    @classmethod
    def from_numpy(cls, array, dtype=None):
        array = np.asarray(array)
        if array.ndim:  # (ascontiguousarray would turn a 0-d scalar into shape [1])
            array = np.ascontiguousarray(array)
        if dtype is None:
            dtype = _TF_OF[array.dtype]
        t = cls.allocate(dtype, array.shape)
        t.numpy()[...] = array.astype(_NP_OF[dtype], copy=False)
        return t

    @property
    def dtype(self):
        return self.fw.TF_TensorType(self.ptr)

    @property
    def shape(self):
        return tuple(self.fw.TF_Dim(self.ptr, i) for i in range(self.fw.TF_NumDims(self.ptr)))

    def numpy(self):
        """A numpy view on the tensor's own storage (no copy; keep `self` alive)."""
        nbytes = self.fw.TF_TensorByteSize(self.ptr)
        np_dtype = np.dtype(_NP_OF[self.dtype])
        if nbytes == 0:
            return np.zeros(self.shape, np_dtype)
        buf = (ctypes.c_char * nbytes).from_address(self.fw.TF_TensorData(self.ptr))
        return np.frombuffer(buf, dtype=np_dtype).reshape(self.shape)

    def __del__(self):
        try:
            if self.owned and self.ptr:
                self.fw.TF_DeleteTensor(self.ptr)
        except Exception:
            pass


class StagedTensor:
    """A feed value already on its way to (or in) device memory: see Session.stage()."""

    def __init__(self, ptr, dtype, shape, nbytes):
        self.fw = framework()
        self.ptr = ptr
        self.dtype = dtype
        self.shape = shape
        self.nbytes = nbytes

    def release(self):
        """Give the device memory back (it belongs to the session's arena, so Session.close()
        releases every staged tensor that is still alive before the device goes away)."""
        if self.ptr:
            self.fw.TF_DeleteTensor(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# ------------------------------------------------------------------ graph
class Output:
    """One output of an Operation (python/framework/ops.py Tensor)."""

    def __init__(self, op, index):
        self.op = op
        self.index = index

    @property
    def name(self):
        return "%s:%d" % (self.op.name, self.index)

    @property
    def dtype(self):
        return framework().TF_OperationOutputType(self._c())

    @property
    def graph(self):
        return self.op.graph

    def _c(self):
        return TF_Output(self.op.ptr, self.index)

    def __repr__(self):
        return "<Output %s dtype=%d>" % (self.name, self.dtype)


class Operation:
    def __init__(self, graph, ptr, inputs, control_inputs, attrs):
        self.graph = graph
        self.ptr = ptr
        self.inputs = list(inputs)
        self.control_inputs = list(control_inputs)
        self.attrs = dict(attrs)
        fw = framework()
        self.name = fw.TF_OperationName(ptr).decode()
        self.type = fw.TF_OperationOpType(ptr).decode()
        self.outputs = [Output(self, i) for i in range(fw.TF_OperationNumOutputs(ptr))]

    def get_attr(self, name):
        return self.attrs[name]

    def __repr__(self):
        return "<Operation %s type=%s>" % (self.name, self.type)


class Graph:
    def __init__(self):
        self.fw = framework()
        self.ptr = self.fw.TF_NewGraph()
        self.operations = []
        self._names = {}
        self.variables = []       # (variable_output, initial_value_output)
        self.shapes = {}          # static shapes recorded by the op constructors (Output.name -> tuple)

    def unique_name(self, base):
        """ops.py Graph.unique_name: base, base_1, base_2, ... skipping names already taken (an
        imported graph may already contain "Const_1")."""
        n = self._names.get(base, 0)
        name = base if n == 0 else "%s_%d" % (base, n)
        while name in self._names and (n > 0 or self._names.get(base, 0) > 0):
            n += 1
            name = "%s_%d" % (base, n)
        self._names[base] = n + 1
        if name != base:
            self._names[name] = self._names.get(name, 0) + 1
        return name

    def create_op(self, op_type, inputs, attrs=None, name=None, control_inputs=(),
                  input_lists=()):
        """ops.py Graph.create_op: inputs are Outputs; attrs maps name -> python value using
        the tagged forms ('type', dt), ('shape', dims), ('tensor', HostTensor), ('ints', [...])."""
        fw = self.fw
        name = self.unique_name(name or op_type)
        desc = fw.TF_NewOperation(self.ptr, op_type.encode(), name.encode())
        for inp in inputs:
            if isinstance(inp, (list, tuple)):
                arr = (TF_Output * len(inp))(*[i._c() for i in inp])
                fw.TF_AddInputList(desc, arr, len(inp))
            else:
                fw.TF_AddInput(desc, inp._c())
        for c in control_inputs:
            fw.TF_AddControlInput(desc, c.ptr)
        status = _Status()
        for k, v in (attrs or {}).items():
            kb = k.encode()
            if isinstance(v, tuple) and v[0] == "type":
                fw.TF_SetAttrType(desc, kb, v[1])
            elif isinstance(v, tuple) and v[0] == "shape":
                dims = (c_int64 * max(len(v[1]), 1))(*v[1])
                fw.TF_SetAttrShape(desc, kb, dims, len(v[1]))
            elif isinstance(v, tuple) and v[0] == "tensor":
                fw.TF_SetAttrTensor(desc, kb, v[1].ptr, status.ptr)
                status.check()
            elif isinstance(v, tuple) and v[0] == "ints":
                arr = (c_int64 * max(len(v[1]), 1))(*v[1])
                fw.TF_SetAttrIntList(desc, kb, arr, len(v[1]))
            elif isinstance(v, bool):
                fw.TF_SetAttrBool(desc, kb, 1 if v else 0)
            elif isinstance(v, int):
                fw.TF_SetAttrInt(desc, kb, v)
            elif isinstance(v, float):
                fw.TF_SetAttrFloat(desc, kb, v)
            elif isinstance(v, str):
                b = v.encode()
                fw.TF_SetAttrString(desc, kb, b, len(b))
            else:
                raise TypeError("unsupported attr %s=%r" % (k, v))
        ptr = fw.TF_FinishOperation(desc, status.ptr)
        status.check()
        flat_inputs = []
        for inp in inputs:
            flat_inputs.extend(inp if isinstance(inp, (list, tuple)) else [inp])
        op = Operation(self, ptr, flat_inputs, control_inputs, attrs or {})
        self.operations.append(op)
        return op

    # ---- GraphDef wire format (framework/graph.proto), hand-written codec on the C++ side
    def as_graph_def(self):
        """The graph as serialized GraphDef bytes (python/framework/ops.py Graph.as_graph_def +
        SerializeToString): what a real TensorFlow 1.0 front-end can import."""
        buf = self.fw.TF_NewBuffer()
        status = _Status()
        self.fw.TF_GraphToGraphDef(self.ptr, buf, status.ptr)
        try:
            status.check()
            b = ctypes.cast(buf, ctypes.POINTER(TF_Buffer)).contents
            return ctypes.string_at(b.data, b.length) if b.length else b""
        finally:
            self.fw.TF_DeleteBuffer(buf)

    def import_graph_def(self, graph_def, name=""):
        """importer.py import_graph_def: add every node of serialized GraphDef bytes to this
        graph (names prefixed with `name/`), e.g. a graph written by real TensorFlow.  Nodes of
        op types this runtime does not register import as opaque nodes (see c_api.h).  Returns
        the new Operations by (prefixed) name."""
        graph_def = bytes(graph_def)
        buf = self.fw.TF_NewBufferFromString(graph_def, len(graph_def))
        opts = self.fw.TF_NewImportGraphDefOptions()
        status = _Status()
        try:
            if name:
                self.fw.TF_ImportGraphDefOptionsSetPrefix(opts, name.encode())
            known = len(self.operations)
            self.fw.TF_GraphImportGraphDef(self.ptr, buf, opts, status.ptr)
            status.check()
        finally:
            self.fw.TF_DeleteImportGraphDefOptions(opts)
            self.fw.TF_DeleteBuffer(buf)
        # wrap the operations the C graph gained (they follow the ones this wrapper created)
        pos = c_size_t(0)
        seen, new_ops = 0, {}
        while True:
            ptr = self.fw.TF_GraphNextOperation(self.ptr, ctypes.byref(pos))
            if not ptr:
                break
            seen += 1
            if seen <= known:
                continue
            op = Operation(self, ptr, [], [], {})
            self.operations.append(op)
            self._names[op.name] = self._names.get(op.name, 0) + 1
            new_ops[op.name] = op
        # second pass (GraphDefs need not be topologically sorted): wire inputs, control inputs
        # and attributes through the C API's introspection calls, record the static shapes that
        # the nodes themselves state (Placeholder / VariableV2 shape attrs, Const values)
        by_ptr = {o.ptr: o for o in self.operations}
        for op in new_ops.values():
            n_in = self.fw.TF_OperationNumInputs(op.ptr)
            for i in range(n_in):
                src = self.fw.TF_OperationInput(TF_Input(op.ptr, i))
                if src.oper in by_ptr and src.index < len(by_ptr[src.oper].outputs):
                    op.inputs.append(by_ptr[src.oper].outputs[src.index])
                elif src.oper in by_ptr:       # output of an opaque node: types unknown
                    op.inputs.append(Output(by_ptr[src.oper], src.index))
            nc = self.fw.TF_OperationNumControlInputs(op.ptr)
            if nc:
                arr = (c_void_p * nc)()
                got = self.fw.TF_OperationGetControlInputs(op.ptr, arr, nc)
                op.control_inputs = [by_ptr[arr[i]] for i in range(got) if arr[i] in by_ptr]
            op.attrs = self._read_attrs(op)
            shape = op.attrs.get("shape")
            if op.type in ("Placeholder", "VariableV2") and isinstance(shape, tuple) and op.outputs:
                self.shapes[op.outputs[0].name] = tuple(shape[1])
            if op.type == "Const" and op.outputs and "_value_shape" in op.attrs:
                self.shapes[op.outputs[0].name] = op.attrs.pop("_value_shape")
        self._infer_shapes(new_ops)
        return new_ops

    def _infer_shapes(self, new_ops):
        """Static shapes of imported hot-path ops, by the same rules the op constructors in ops.py
        apply (the reference's shape functions, framework/common_shape_fns.cc): gradient
        construction needs them.  Anything unknown simply stays unknown."""
        def windowed(size, filt, stride, padding):
            return (size - filt + stride) // stride if padding == "VALID" else (size + stride - 1) // stride

        def ints(op, key):
            v = op.attrs.get(key)
            return v[1] if isinstance(v, tuple) else None

        pending = [op for op in new_ops.values() if op.outputs]
        for _ in range(len(pending) + 1):       # GraphDefs need not be sorted: iterate to a fixed point
            progress = False
            for op in pending:
                if op.outputs[0].name in self.shapes and op.type != "SoftmaxCrossEntropyWithLogits":
                    continue
                ins = [self.shapes.get(i.name) for i in op.inputs]
                out = None
                t = op.type
                if t in ("Identity", "Relu", "Softmax", "LogSoftmax", "Cast", "BiasAdd", "ReluGrad"):
                    out = ins[0] if ins else None
                elif t in ("Add", "Mul") and len(ins) == 2 and None not in ins:
                    a, b = ins
                    out = a if int(np.prod(a or (1,))) != 1 or len(a) >= len(b) else b
                    if int(np.prod(a or (1,))) == 1 and int(np.prod(b or (1,))) != 1:
                        out = b
                elif t == "MatMul" and len(ins) == 2 and None not in ins:
                    a, b = ins
                    out = (a[1] if op.attrs.get("transpose_a") else a[0],
                           b[0] if op.attrs.get("transpose_b") else b[1])
                elif t == "BatchMatMul" and len(ins) == 2 and None not in ins:
                    a, b = ins
                    out = tuple(a[:-2]) + (a[-1] if op.attrs.get("adj_x") else a[-2],
                                           b[-2] if op.attrs.get("adj_y") else b[-1])
                elif t in ("Conv2D", "MaxPool") and ins and ins[0] is not None:
                    nchw = op.attrs.get("data_format", "NHWC") == "NCHW"
                    h, w, c = (2, 3, 1) if nchw else (1, 2, 3)
                    st, pad, x = ints(op, "strides"), op.attrs.get("padding"), ins[0]
                    if t == "Conv2D" and len(ins) == 2 and ins[1] is not None and st:
                        f = ins[1]
                        oh, ow, oc = windowed(x[h], f[0], st[h], pad), windowed(x[w], f[1], st[w], pad), f[3]
                    elif t == "MaxPool" and st and ints(op, "ksize"):
                        k = ints(op, "ksize")
                        oh, ow, oc = windowed(x[h], k[h], st[h], pad), windowed(x[w], k[w], st[w], pad), x[c]
                    else:
                        continue
                    out = (x[0], oc, oh, ow) if nchw else (x[0], oh, ow, oc)
                elif t == "Reshape" and len(op.inputs) == 2 and ins[0] is not None:
                    target = op.inputs[1].op.attrs.get("_const_ints")
                    if target is not None:
                        n = int(np.prod(ins[0], dtype=np.int64))
                        known = int(np.prod([d for d in target if d >= 0], dtype=np.int64)) or 1
                        out = tuple(d if d >= 0 else n // known for d in target)
                elif t == "SoftmaxCrossEntropyWithLogits" and ins and ins[0] is not None:
                    if op.outputs[1].name not in self.shapes:
                        self.shapes[op.outputs[0].name] = (ins[0][0],)
                        self.shapes[op.outputs[1].name] = tuple(ins[0])
                        progress = True
                    continue
                elif t == "Mean" and ins and ins[0] is not None:
                    out = ()     # the runtime's Mean reduces over all dimensions
                elif t == "ArgMax" and len(op.inputs) == 2 and ins[0] is not None:
                    axis = op.inputs[1].op.attrs.get("_const_ints")
                    if axis:
                        ax = axis[0] % len(ins[0])
                        out = tuple(d for i, d in enumerate(ins[0]) if i != ax)
                if out is not None:
                    self.shapes[op.outputs[0].name] = tuple(int(d) for d in out)
                    progress = True
            if not progress:
                break

    def _read_attrs(self, op):
        """The op's attributes in the tagged form create_op takes (so gradient functions can copy
        them): bool / int / float / str, ('type', dt), ('shape', dims), ('ints', [...]).  Attributes
        the runtime keeps as opaque bytes, and tensors, are left out."""
        fw, st = self.fw, _Status()
        names = _free_cstr_list(fw.B200TF_OperationAttrNames(op.ptr))
        out = {}
        for name in names:
            nb = name.encode()
            m = fw.TF_OperationGetAttrMetadata(op.ptr, nb, st.ptr)
            st.check()
            if m.is_list:
                if m.type == 1:  # TF_ATTR_INT
                    vals = (c_int64 * max(m.list_size, 1))()
                    fw.TF_OperationGetAttrIntList(op.ptr, nb, vals, m.list_size, st.ptr)
                    out[name] = ("ints", [int(v) for v in vals[:m.list_size]])
                continue
            if m.type == 0:      # string
                buf = ctypes.create_string_buffer(max(m.total_size, 1))
                fw.TF_OperationGetAttrString(op.ptr, nb, buf, m.total_size, st.ptr)
                out[name] = buf.raw[:m.total_size].decode("utf-8", "replace")
            elif m.type == 1:
                v = c_int64()
                fw.TF_OperationGetAttrInt(op.ptr, nb, ctypes.byref(v), st.ptr)
                out[name] = int(v.value)
            elif m.type == 2:
                v = ctypes.c_float()
                fw.TF_OperationGetAttrFloat(op.ptr, nb, ctypes.byref(v), st.ptr)
                out[name] = float(v.value)
            elif m.type == 3:
                v = ctypes.c_ubyte()
                fw.TF_OperationGetAttrBool(op.ptr, nb, ctypes.byref(v), st.ptr)
                out[name] = bool(v.value)
            elif m.type == 4:
                v = c_int()
                fw.TF_OperationGetAttrType(op.ptr, nb, ctypes.byref(v), st.ptr)
                out[name] = ("type", int(v.value))
            elif m.type == 5 and m.total_size >= 0:
                dims = (c_int64 * max(m.total_size, 1))()
                fw.TF_OperationGetAttrShape(op.ptr, nb, dims, m.total_size, st.ptr)
                out[name] = ("shape", [int(d) for d in dims[:m.total_size]])
            elif m.type == 6 and name == "value":
                t = c_void_p()
                fw.TF_OperationGetAttrTensor(op.ptr, nb, ctypes.byref(t), st.ptr)
                if t.value:
                    ht = HostTensor(t.value)
                    out["_value_shape"] = ht.shape
                    if ht.dtype in (TF_INT32, TF_INT64) and int(np.prod(ht.shape, dtype=np.int64)) <= 16:
                        out["_const_ints"] = [int(v) for v in np.array(ht.numpy()).ravel()]
            st.check()
        return out

    def get_operation_by_name(self, name):
        for op in self.operations:
            if op.name == name:
                return op
        raise KeyError("no operation named %r" % name)

    def get_tensor_by_name(self, name):
        op_name, _, idx = name.partition(":")
        op = self.get_operation_by_name(op_name)
        return op.outputs[int(idx or 0)]

    def __del__(self):
        try:
            self.fw.TF_DeleteGraph(self.ptr)
        except Exception:
            pass


def graph_def_to_text(graph_def):
    """One line per node of serialized GraphDef bytes (name, op, device, inputs, attr summaries):
    a dump for eyes and tests, produced by the C++ wire reader."""
    fw = framework()
    status = _Status()
    graph_def = bytes(graph_def)
    p = fw.B200TF_GraphDefToText(graph_def, len(graph_def), status.ptr)
    status.check()
    try:
        return ctypes.string_at(p).decode("utf-8", "replace")
    finally:
        ctypes.CDLL(None).free(c_void_p(p))


# ------------------------------------------------------------------ session
class Session:
    """session.py BaseSession: run(fetches, feed_dict).  Feeds are host arrays, fetches come back
    as numpy arrays; one device sync per run."""

    def __init__(self, graph, gpu=0, memory_limit_bytes=0, collective_comm=None, num_replicas=1):
        self.fw = framework()
        self.graph = graph
        opts = self.fw.TF_NewSessionOptions()
        self.fw.B200TF_SetGpuDevice(opts, gpu)
        if memory_limit_bytes:
            self.fw.B200TF_SetGpuMemoryLimit(opts, memory_limit_bytes)
        if collective_comm:
            self.fw.B200TF_SetCollective(opts, collective_comm, num_replicas)
        status = _Status()
        self.ptr = self.fw.TF_NewSession(graph.ptr, opts, status.ptr)
        self.fw.TF_DeleteSessionOptions(opts)
        status.check()
        self._status = _Status()
        self._staged = weakref.WeakSet()

    def run(self, fetches, feed_dict=None, as_host_tensors=False):
        single = not isinstance(fetches, (list, tuple))
        fetch_list = [fetches] if single else list(fetches)
        outs = [f for f in fetch_list if isinstance(f, Output)]
        targets = [f for f in fetch_list if isinstance(f, Operation)]
        feed_dict = feed_dict or {}
        keep = []
        feed_outputs, feed_ptrs = [], []
        for k, v in feed_dict.items():
            if isinstance(v, StagedTensor) and not v.ptr:
                raise ValueError("staged tensor fed to %s was already released" % k.name)
            if not isinstance(v, (HostTensor, StagedTensor)):
                v = HostTensor.from_numpy(np.asarray(v), k.dtype)
            keep.append(v)
            feed_outputs.append(k._c())
            feed_ptrs.append(v.ptr)
        n_in, n_out, n_t = len(feed_outputs), len(outs), len(targets)
        c_in = (TF_Output * max(n_in, 1))(*feed_outputs)
        c_in_vals = (c_void_p * max(n_in, 1))(*feed_ptrs)
        c_out = (TF_Output * max(n_out, 1))(*[o._c() for o in outs])
        c_out_vals = (c_void_p * max(n_out, 1))()
        c_targets = (c_void_p * max(n_t, 1))(*[t.ptr for t in targets])
        self.fw.TF_SessionRun(self.ptr, None, c_in, c_in_vals, n_in, c_out, c_out_vals, n_out,
                              c_targets, n_t, None, self._status.ptr)
        self._status.check()
        results = []
        it = iter(range(n_out))
        for f in fetch_list:
            if isinstance(f, Output):
                t = HostTensor(c_out_vals[next(it)])
                results.append(t if as_host_tensors else np.array(t.numpy()))
            else:
                results.append(None)
        return results[0] if single else results

    def stage(self, value, dtype=None):
        """Start copying a feed value to the device and return at once (input prefetch).

        `value` is a HostTensor (pinned; its buffer must not be rewritten until the run that
        consumes the staged tensor has returned) or anything numpy can convert.  The returned
        StagedTensor is accepted by run() as a feed value and is used without another copy, so
        staging step i+1 before running step i hides the PCIe transfer behind the kernels.
        """
        if not isinstance(value, HostTensor):
            value = HostTensor.from_numpy(np.asarray(value), dtype)
        ptr = self.fw.B200TF_SessionStageTensor(self.ptr, value.ptr, self._status.ptr)
        self._status.check()
        st = StagedTensor(ptr, value.dtype, value.shape, self.fw.TF_TensorByteSize(value.ptr))
        self._staged.add(st)
        return st

    def stream(self):
        """The CUstream handle (int) all kernels of this session run on."""
        return self.fw.B200TF_SessionStream(self.ptr)

    def last_run_stats(self):
        s = RunStats()
        self.fw.B200TF_SessionLastRunStats(self.ptr, ctypes.byref(s))
        return {"nodes_executed": s.nodes_executed, "kernels_launched": s.kernels_launched,
                "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes,
                "host_enqueue_us": s.host_enqueue_us, "host_total_us": s.host_total_us}

    def close(self):
        if self.ptr:
            for staged in list(self._staged):
                staged.release()
            st = _Status()
            self.fw.TF_CloseSession(self.ptr, st.ptr)
            self.fw.TF_DeleteSession(self.ptr, st.ptr)
            self.ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
