"""The BASELINE.json workloads (configs[1..3]) as (a) graphs built through the product's Python
front-end and (b) the same training step chained on the CPU oracle.

TEST / BENCH INFRASTRUCTURE.  Shared by tests/test_session_fullsize_gpu.py (full-size Session
parity), bench.py (the `parity` block computed outside the timed region, the cpu_baseline leg and
--impl reference) so that the graph that is timed is the graph that is checked.  The product
package never imports this module; the `reference_*` halves are the only code here that touches
the oracle.

Synthetic inputs per SURVEY.md 8d: activations U(-1,1), weights N(0, 1/sqrt(fan_in)), biases 0.1,
one-hot labels, fixed seeds; `seed` selects a replica's data shard, initial weights are the same
on every replica; bf16 configs = the same fp32 data truncated to bf16.
"""
from collections import OrderedDict

import numpy as np

LR = 0.01


# ------------------------------------------------------------------ bf16 helpers (numpy)
def bf16_truncate(a):
    """fp32 -> the bf16-representable fp32 below it in magnitude (the reference's conversion,
    core/framework/bfloat16.cc:20-50: keep the high 16 bits)."""
    a = np.ascontiguousarray(a, np.float32)
    return (a.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def bf16_round(a):
    """fp32 -> nearest-even bf16, as fp32 (what a kernel's fp32 accumulator -> bf16 store does)."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def bf16_bits(a):
    """bf16-representable fp32 -> uint16 bit patterns."""
    return (np.ascontiguousarray(a, np.float32).view(np.uint32) >> 16).astype(np.uint16)


def bf16_from_bits(bits):
    return (np.ascontiguousarray(bits, np.uint16).astype(np.uint32) << 16).view(np.float32)


def tf32_truncate(a):
    """fp32 -> the value tcgen05 kind::tf32 multiplies with: the low 13 mantissa bits are ignored
    (DESIGN.md section 3).  Sums stay fp32."""
    a = np.ascontiguousarray(a, np.float32)
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


class KernelRounding:
    """The oracle evaluated under the tensor-core kernels' documented INPUT rounding: MatMul and
    the implicit-GEMM convolutions read their fp32 operands as TF32 (truncation), accumulate in
    fp32.  First-layer convolutions (C_in <= 4) run on the CUDA cores in IEEE fp32 and are left
    exact.  Everything else is forwarded to the oracle unchanged.

    Why it exists: a ReLU network's gradient is discontinuous in its pre-activations -- a forward
    pass that differs by TF32 rounding (5e-4) flips a fraction f ~ 5e-4 of the ReLU masks, which
    moves the early layers' gradients by ~sqrt(f) ~ 2 % in Frobenius norm whatever the backward
    kernels do (tests/test_workloads_cpu.py reproduces the figure on the CPU alone).  Comparing
    with the oracle under the same input rounding removes that effect and checks the kernels
    themselves: summation, masks, layout, scaling."""

    def __init__(self, o):
        self._o = o

    def __getattr__(self, name):
        return getattr(self._o, name)

    def matmul(self, a, b, transpose_a=False, transpose_b=False):
        return self._o.matmul(tf32_truncate(a), tf32_truncate(b), transpose_a, transpose_b)

    def conv2d(self, x, f, strides, padding):
        if x.shape[-1] <= 4:
            return self._o.conv2d(x, f, strides, padding)
        return self._o.conv2d(tf32_truncate(x), tf32_truncate(f), strides, padding)

    def conv2d_backprop_input(self, in_shape, f, dy, strides, padding):
        if in_shape[-1] <= 4:
            return self._o.conv2d_backprop_input(in_shape, f, dy, strides, padding)
        return self._o.conv2d_backprop_input(in_shape, tf32_truncate(f), tf32_truncate(dy), strides,
                                             padding)

    def conv2d_backprop_filter(self, x, filter_shape, dy, strides, padding):
        if x.shape[-1] <= 4:
            return self._o.conv2d_backprop_filter(x, filter_shape, dy, strides, padding)
        return self._o.conv2d_backprop_filter(tf32_truncate(x), filter_shape, tf32_truncate(dy),
                                              strides, padding)


def bucket_kw():
    """B200TF_BUCKET_BYTES=none|<bytes>: gradient all-reduce bucket size (default: optimizer's)."""
    import os
    v = os.environ.get("B200TF_BUCKET_BYTES")
    if not v:
        return {}
    return {"bucket_bytes": None if v == "none" else int(v)}


class Built:
    """What Workload.build() returns: handles into the graph."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class Workload:
    name = ""
    batch = 0
    dtype = "f32"   # storage type of activations / weights / gradients

    # ---- data
    def data(self, seed):
        raise NotImplementedError

    def init_params(self):
        raise NotImplementedError

    # ---- graph (product API only)
    def tower(self, tf, V, inp, lab):
        raise NotImplementedError

    def input_shapes(self):
        raise NotImplementedError

    def build(self, num_replicas=1, seed=1234, resident=True):
        """Two towers over shared variables: `resident` reads x / labels from device variables
        (inputs already in HBM), `fed` from placeholders (host-fed)."""
        from simple_tensorflow_b200 import ops as tf
        x, labels = self.data(seed)
        params = self.init_params()
        tdt = tf.bfloat16 if self.dtype == "bf16" else tf.float32
        tf.reset_default_graph()
        V = OrderedDict((n, tf.Variable(a, dtype=tdt, name=n)) for n, a in params.items())
        train_vars = list(V.values())

        def finish(loss, tag):
            opt = tf.GradientDescentOptimizer(LR)
            gv = opt.compute_gradients(loss, train_vars)
            train = opt.apply_gradients(gv, name=tag + "/train", num_replicas=num_replicas,
                                        **bucket_kw())
            return loss, train, opt

        xs, ls = self.input_shapes()
        res = None
        if resident:
            x_res = tf.Variable(x, dtype=tdt, name="x_resident")
            l_res = tf.Variable(labels, dtype=tdt, name="labels_resident")
            res = finish(self.tower(tf, V, x_res.ref, l_res.ref, "resident"), "resident")
        xp = tf.placeholder(tdt, list(xs), "x")
        lp = tf.placeholder(tdt, list(ls), "labels")
        fed = finish(self.tower(tf, V, xp, lp, "fed"), "fed")
        # the gradients ApplyGradientDescent consumes (after the replica average when N > 1)
        applied = [(v.op.name, g) for g, v in fed[2].applied_gradients]
        return Built(tf=tf, x=x, labels=labels, params=params, V=V, xp=xp, lp=lp,
                     resident=None if res is None else (res[0], res[1]), fed=(fed[0], fed[1]),
                     applied_grads=applied, tdt=tdt)

    def host_tensor(self, array):
        """A pinned host tensor of this workload's storage type holding `array` (fp32 values that
        are bf16-representable for the bf16 workloads)."""
        from simple_tensorflow_b200 import client, ops as tf
        if self.dtype == "bf16":
            t = client.HostTensor.allocate(tf.bfloat16, array.shape)
            t.numpy()[...] = bf16_bits(array)
            return t
        return client.HostTensor.from_numpy(np.ascontiguousarray(array, np.float32))

    def to_f32(self, fetched):
        """A fetched value of the storage type -> fp32 numpy."""
        a = np.asarray(fetched)
        if a.dtype == np.uint16:
            return bf16_from_bits(a)
        return a.astype(np.float32, copy=False)

    # ---- CPU oracle
    def reference(self, o, x, labels, params):
        """-> (mean loss, {name: gradient of the mean loss}) on the CPU oracle."""
        raise NotImplementedError

    def reference_step(self, o, x, labels, params):
        """One full training step on the CPU: loss, gradients and the SGD update (what the
        cpu_baseline / --impl reference legs time).  Updates `params` in place."""
        loss, grads = self.reference(o, x, labels, params)
        for n in params:
            params[n] = o.apply_gradient_descent(params[n], LR, grads[n])
        return loss

    def _q(self, a):
        """Storage rounding of an op output in the oracle chain (identity for fp32)."""
        return bf16_round(a) if self.dtype == "bf16" else a


class MLP(Workload):
    """BASELINE configs[1] (fp32) / configs[3] per replica (bf16): 3 x 1024 MLP, batch 4096,
    softmax cross-entropy over 1024 classes, fwd + bwd + SGD."""

    def __init__(self, dtype="f32", batch=4096, width=1024, layers=3):
        self.dtype, self.batch, self.width, self.layers = dtype, batch, width, layers
        self.name = "mlp" if dtype == "f32" else "mlp_bf16"
        # every GEMM of the step is 2*B*W*W: 3 forward, 3 dW, 2 dX (the input is data)
        self.gemm_flops = 2.0 * batch * width * width
        self.flops_per_step = (3 * layers - 1) * self.gemm_flops
        self.describe = ("mlp-%dx%d batch %d/replica %s fwd+bwd+sgd (BASELINE configs[%d]); "
                         "Session.Run([loss, train_op])" %
                         (layers, width, batch, "fp32" if dtype == "f32" else "bf16",
                          1 if dtype == "f32" else 3))

    def input_shapes(self):
        return (self.batch, self.width), (self.batch, self.width)

    def data(self, seed):
        rng = np.random.RandomState(seed)
        x = rng.uniform(-1, 1, (self.batch, self.width)).astype(np.float32)
        labels = np.zeros((self.batch, self.width), np.float32)
        labels[np.arange(self.batch), rng.randint(0, self.width, self.batch)] = 1.0
        if self.dtype == "bf16":
            x = bf16_truncate(x)
        return x, labels

    def init_params(self):
        rng = np.random.RandomState(4321)
        p = OrderedDict()
        for i in range(self.layers):
            w = (rng.randn(self.width, self.width) / np.sqrt(self.width)).astype(np.float32)
            p["W%d" % i] = bf16_truncate(w) if self.dtype == "bf16" else w
        for i in range(self.layers):
            b = np.full(self.width, 0.1, np.float32)
            p["b%d" % i] = bf16_truncate(b) if self.dtype == "bf16" else b
        return p

    def tower(self, tf, V, inp, lab, tag):
        h = inp
        for i in range(self.layers):
            h = tf.bias_add(tf.matmul(h, V["W%d" % i], name="%s/fc%d" % (tag, i)), V["b%d" % i])
            if i < self.layers - 1:
                h = tf.relu(h)
        return tf.reduce_mean(tf.softmax_cross_entropy_with_logits(h, lab), name=tag + "/loss")

    def reference(self, o, x, labels, params):
        # bf16: storage rounding where the kernels round -- once per FUSED op (MatMul+BiasAdd(+Relu)
        # and MatMul+ReluGrad each round their fp32 accumulator once; xent scales its backprop by
        # 1/batch in fp32 before the store)
        q = self._q
        L = self.layers
        acts = [x]
        for i in range(L):
            pre = q(o.bias_add(o.matmul(acts[-1], params["W%d" % i]), params["b%d" % i]))
            acts.append(o.relu(pre) if i < L - 1 else pre)
        lvec, bp = o.softmax_xent(acts[-1], labels)
        g = q(bp * np.float32(1.0 / x.shape[0]))
        grads = {}
        for i in reversed(range(L)):
            grads["b%d" % i] = q(o.bias_add_grad(g))
            grads["W%d" % i] = q(o.matmul(acts[i], g, True, False))
            if i > 0:
                g = q(o.relu_grad(o.matmul(g, params["W%d" % i], False, True), acts[i]))
        return float(q(lvec).mean()), grads


class LeNet(Workload):
    """BASELINE configs[2] / SURVEY 8d C3: conv5x5x1x32 SAME -> relu -> pool2 -> conv5x5x32x64 SAME
    -> relu -> pool2 -> fc 3136x1024 + relu -> fc 1024x10 -> xent; batch 512, NHWC fp32."""

    SHAPES = OrderedDict(w1=(5, 5, 1, 32), w2=(5, 5, 32, 64), w3=(3136, 1024), w4=(1024, 10))

    def __init__(self, batch=512):
        self.batch, self.dtype, self.name = batch, "f32", "lenet"
        B = batch
        conv1 = 2.0 * B * 28 * 28 * 25 * 1 * 32
        conv2 = 2.0 * B * 14 * 14 * 25 * 32 * 64
        fc1 = 2.0 * B * 3136 * 1024
        fc2 = 2.0 * B * 1024 * 10
        # forward + filter gradients everywhere + input gradients except into the data
        self.flops_per_step = 2 * conv1 + 3 * conv2 + 3 * fc1 + 3 * fc2
        self.describe = ("lenet-5 batch %d/replica NHWC fp32 fwd+bwd+sgd (BASELINE configs[2]); "
                         "Session.Run([loss, train_op])" % batch)

    def input_shapes(self):
        return (self.batch, 28, 28, 1), (self.batch, 10)

    def data(self, seed):
        rng = np.random.RandomState(seed)
        x = rng.uniform(-1, 1, (self.batch, 28, 28, 1)).astype(np.float32)
        labels = np.zeros((self.batch, 10), np.float32)
        labels[np.arange(self.batch), rng.randint(0, 10, self.batch)] = 1.0
        return x, labels

    def init_params(self):
        rng = np.random.RandomState(4321)
        p = OrderedDict()
        for n, shp in self.SHAPES.items():
            fan_in = int(np.prod(shp[:-1]))
            p[n] = (rng.randn(*shp) / np.sqrt(fan_in)).astype(np.float32)
            p["b" + n[1]] = np.full(shp[-1], 0.1, np.float32)
        return p

    def tower(self, tf, V, inp, lab, tag):
        B = self.batch
        c1 = tf.relu(tf.bias_add(tf.conv2d(inp, V["w1"], [1, 1, 1, 1], "SAME"), V["b1"]))
        p1 = tf.max_pool(c1, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
        c2 = tf.relu(tf.bias_add(tf.conv2d(p1, V["w2"], [1, 1, 1, 1], "SAME"), V["b2"]))
        p2 = tf.max_pool(c2, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
        flat = tf.reshape(p2, [B, 3136])
        f1 = tf.relu(tf.bias_add(tf.matmul(flat, V["w3"]), V["b3"]))
        logits = tf.bias_add(tf.matmul(f1, V["w4"]), V["b4"])
        return tf.reduce_mean(tf.softmax_cross_entropy_with_logits(logits, lab), name=tag + "/loss")

    def reference(self, o, x, labels, P):
        B = x.shape[0]
        a1 = o.relu(o.bias_add(o.conv2d(x, P["w1"], (1, 1), "SAME"), P["b1"]))
        q1 = o.max_pool(a1, (2, 2), (2, 2), "SAME")
        a2 = o.relu(o.bias_add(o.conv2d(q1, P["w2"], (1, 1), "SAME"), P["b2"]))
        q2 = o.max_pool(a2, (2, 2), (2, 2), "SAME")
        fl = q2.reshape(B, -1)
        g1 = o.relu(o.bias_add(o.matmul(fl, P["w3"]), P["b3"]))
        lg = o.bias_add(o.matmul(g1, P["w4"]), P["b4"])
        lvec, bp = o.softmax_xent(lg, labels)
        d = bp * np.float32(1.0 / B)
        G = {"b4": o.bias_add_grad(d), "w4": o.matmul(g1, d, True, False)}
        d = o.relu_grad(o.matmul(d, P["w4"], False, True), g1)
        G["b3"] = o.bias_add_grad(d)
        G["w3"] = o.matmul(fl, d, True, False)
        d = o.matmul(d, P["w3"], False, True).reshape(q2.shape)
        d = o.relu_grad(o.max_pool_grad(a2, d, (2, 2), (2, 2), "SAME"), a2)
        G["b2"] = o.bias_add_grad(d)
        G["w2"] = o.conv2d_backprop_filter(q1, P["w2"].shape, d, (1, 1), "SAME")
        d = o.conv2d_backprop_input(q1.shape, P["w2"], d, (1, 1), "SAME")
        d = o.relu_grad(o.max_pool_grad(a1, d, (2, 2), (2, 2), "SAME"), a1)
        G["b1"] = o.bias_add_grad(d)
        G["w1"] = o.conv2d_backprop_filter(x, P["w1"].shape, d, (1, 1), "SAME")
        return float(lvec.mean()), G


def get(name):
    if name == "mlp":
        return MLP("f32")
    if name == "mlp_bf16":
        return MLP("bf16")
    if name == "lenet":
        return LeNet()
    raise KeyError(name)


# ------------------------------------------------------------------ parity of a built session
def rel_fro(got, ref):
    ref = np.asarray(ref, np.float64)
    d = np.linalg.norm((np.asarray(got, np.float64) - ref).ravel())
    n = np.linalg.norm(ref.ravel())
    return float(d / n) if n > 0 else float(d)


def check_parity(w, B, sess, o, world=1, rank=0, seeds=None, tol=1e-2):
    """Step-1 parity of the FULL-SIZE fed tower against the CPU oracle, outside any timed region.

    Run on every rank (the gradient exchange is collective); the comparison is done where it is
    called (rank 0 uses it).  Checks, in this order, without and then with the update applied:
      loss        rank-local mean loss vs the oracle on this rank's batch
      grads       the gradients ApplyGradientDescent consumes (replica-averaged for N > 1) vs the
                  oracle's gradient of the GLOBAL batch (mean over the replicas' batches)
      weights     variables after one Session.Run([loss, train]) vs  w0 - lr * oracle gradient
    Errors are relative Frobenius norms; `ok` = all <= tol (north_star: 1e-2 relative fp32).
    """
    seeds = seeds or [1234 + r for r in range(world)]
    hx, hl = w.host_tensor(B.x), w.host_tensor(B.labels)
    feed = {B.xp: hx, B.lp: hl}
    names = [n for n, _ in B.applied_grads]
    vals = sess.run([B.fed[0]] + [g for _, g in B.applied_grads], feed)
    got_loss = float(np.asarray(w.to_f32(vals[0])).reshape(-1)[0])
    got_grads = {n: w.to_f32(v) for n, v in zip(names, vals[1:])}
    sess.run([B.fed[0], B.fed[1]], feed)
    got_w = {n: w.to_f32(a) for n, a in zip(B.V, sess.run([v.ref for v in B.V.values()]))}

    if rank != 0 and world > 1:
        # The GPU runs above are collective (every rank takes part); the CPU oracle of the global
        # batch is evaluated on rank 0 only -- N ranks x all host threads each would oversubscribe
        # the box's CPU quota N-fold (8 ranks: minutes instead of seconds).
        return None

    def global_reference(oracle_like):
        loss_local, grads = None, None
        for r, seed in enumerate(seeds):
            x, labels = w.data(seed)
            loss_r, g_r = w.reference(oracle_like, x, labels, B.params)
            if r == rank:
                loss_local = loss_r
            if grads is None:
                grads = {n: np.asarray(g, np.float64) for n, g in g_r.items()}
            else:
                for n in grads:
                    grads[n] += g_r[n]
        return loss_local, {n: (g / len(seeds)).astype(np.float32) for n, g in grads.items()}

    ref_loss_local, ref_grads = global_reference(o)
    grad_err = {n: rel_fro(got_grads[n], ref_grads[n]) for n in names}
    # the same comparison with the oracle under the kernels' input rounding (KernelRounding): for
    # fp32 graphs the TF32 operand truncation; the bf16 chain already models its storage rounding
    if w.dtype == "f32":
        _, kr_grads = global_reference(KernelRounding(o))
        grad_err_kr = {n: rel_fro(got_grads[n], kr_grads[n]) for n in names}
    else:
        grad_err_kr = dict(grad_err)
    w_err = {}
    for n in B.V:
        ref_w = B.params[n] - np.float32(LR) * ref_grads[n]
        w_err[n] = rel_fro(got_w[n], w._q(ref_w) if w.dtype == "bf16" else ref_w)
    loss_err = abs(got_loss - ref_loss_local) / max(abs(ref_loss_local), 1e-30)
    worst_g = max(grad_err.values())
    worst_g_kr = max(grad_err_kr.values())
    worst_w = max(w_err.values())
    return {"ok": bool(loss_err <= tol and worst_w <= tol and worst_g_kr <= tol), "tol": tol,
            "loss": got_loss, "loss_oracle": ref_loss_local, "loss_rel_err": loss_err,
            "weight_rel_err_max": worst_w,
            "grad_rel_err_max_same_input_rounding": worst_g_kr,
            "grad_rel_err_max_vs_exact_fp32": worst_g,
            "grad_rel_err_same_input_rounding": grad_err_kr,
            "grad_rel_err_vs_exact_fp32": grad_err,
            "what": "step-1 loss, updated weights and gradients (as consumed by "
                    "ApplyGradientDescent%s) of the full-size graph vs the CPU oracle%s; relative "
                    "Frobenius norms.  ok = loss, weights (vs the exact fp32 oracle) and gradients "
                    "(vs the oracle under the kernels' documented input rounding: %s) all <= tol.  "
                    "The gradients vs the exact fp32 oracle are reported too: ReLU masks flipped by "
                    "the forward pass's rounding move early-layer gradients by ~sqrt(flip fraction) "
                    "whatever the backward kernels do (tests/workloads.py::KernelRounding)."
                    % (", replica-averaged" if world > 1 else "",
                       " on the global batch of %d replicas" % world if world > 1 else "",
                       "TF32 operand truncation" if w.dtype == "f32" else
                       "bf16 storage, one rounding per fused op")}
