#!/usr/bin/env python3
"""bench.py -- Session.Run samples/sec of the B200 op-kernel layer (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, see below)

Workload at N=1 (BASELINE.json configs[1], SURVEY 8d "C2"): 3-layer MLP 1024-1024-1024, batch
4096, fp32 graph, forward + backward + SGD update, softmax cross-entropy over 1024 classes.
One "step" = one Session.Run([loss, train_op]) through the C API (TF_SessionRun).
N>1: one process per GPU (torchrun), one graph replica per rank, weak scaling (4096 samples per
rank), gradients averaged by ONE fused NCCL all-reduce per step (B200AllReduceN).

  value : samples/s with inputs resident in HBM (x / labels live in device variables).
  e2e   : the same step fed from pinned HOST buffers through TF_SessionRun: H2D of x+labels and
          D2H of the loss inside the timed region.
  roofline     : the tcgen05 GEMM (dominant kernel), device time per launch measured with CUDA
                 events on the session's stream (b200_profile_begin/end) in a second pass.
  cpu_baseline : the CPU oracle (oracle/oracle.c, a restatement of the reference's Eigen path;
                 the reference itself cannot be built offline) on the box's host cores.
--impl reference: times that same CPU port with all host threads (the reference's own CPU
implementation is unbuildable here: no bazel / protoc / Eigen).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Session.Run samples/sec (MLP-1024 & LeNet) at 1/2/4/8 B200 vs Eigen CPU"
BATCH, WIDTH, LAYERS, LR = 4096, 1024, 3, 0.01
GEMM_FLOPS = 2.0 * BATCH * WIDTH * WIDTH  # every GEMM of the step: 8.59 GFLOP
STEP_GEMMS = 8                            # 3 fwd + 3 dW + 2 dX (input is data)


def synthetic(seed):
    """SURVEY 8d inputs: activations U(-1,1), weights N(0, 1/sqrt(fan_in)), biases 0.1,
    one-hot labels, fixed seed.  `seed` selects the replica's data shard; the initial weights
    are the same on every replica."""
    rng = np.random.RandomState(seed)
    x = rng.uniform(-1, 1, (BATCH, WIDTH)).astype(np.float32)
    labels = np.zeros((BATCH, WIDTH), np.float32)
    labels[np.arange(BATCH), rng.randint(0, WIDTH, BATCH)] = 1.0
    rng = np.random.RandomState(4321)
    ws = [(rng.randn(WIDTH, WIDTH) / np.sqrt(WIDTH)).astype(np.float32) for _ in range(LAYERS)]
    bs = [np.full(WIDTH, 0.1, np.float32) for _ in range(LAYERS)]
    return x, labels, ws, bs


def bucket_kw():
    """B200TF_BUCKET_BYTES=none|<bytes>: gradient all-reduce bucket size (default: optimizer's)."""
    v = os.environ.get("B200TF_BUCKET_BYTES")
    if not v:
        return {}
    return {"bucket_bytes": None if v == "none" else int(v)}


# =================================================================================== CPU arm
def _cgroup_cpu_quota():
    """CPUs this container may use per the cgroup bandwidth controller (None = unlimited)."""
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]  # cgroup v2
        if quota != "max":
            return max(1, -(-int(quota) // int(period)))
        return None
    except Exception:
        pass
    try:  # cgroup v1
        quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if quota > 0 and period > 0:
            return max(1, -(-quota // period))
    except Exception:
        pass
    return None


def host_threads():
    """Threads of the CPU arms = the CPUs this process can really use: the schedulable CPUs (what
    the reference's NumSchedulableCPUs returns, platform/posix/port.cc:50-55), capped by the
    physical core count and by the container's cgroup CPU quota.  Measured on the GPU box
    (128 logical / 64 physical CPUs, cpu.max = 16 CPUs): 16 threads 40.2 K samples/s, 64 threads
    21.0 K, 128 threads 2.7 K -- oversubscribing the quota only throttles, and the baseline
    should be the CPU's best.  B200TF_HOST_THREADS overrides."""
    if os.environ.get("B200TF_HOST_THREADS"):
        return max(1, int(os.environ["B200TF_HOST_THREADS"]))
    n = len(os.sched_getaffinity(0))
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys)
    except Exception:
        pass
    quota = _cgroup_cpu_quota()
    if quota:
        n = min(n, quota)
    return max(1, n)


def cpu_step_fn():
    """One full MLP training step on the CPU oracle (test infrastructure used as the baseline)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bind as o
    o.set_num_threads(host_threads())
    x, labels, ws, bs = synthetic(1234)

    def step():
        acts = [x]
        for i in range(LAYERS):
            pre = o.bias_add(o.matmul(acts[-1], ws[i]), bs[i])
            acts.append(o.relu(pre) if i < LAYERS - 1 else pre)
        lvec, bp = o.softmax_xent(acts[-1], labels)
        g = bp * np.float32(1.0 / BATCH)
        for i in reversed(range(LAYERS)):
            db = o.bias_add_grad(g)
            dw = o.matmul(acts[i], g, True, False)
            if i > 0:
                g = o.relu_grad(o.matmul(g, ws[i], False, True), acts[i])
            ws[i] = o.apply_gradient_descent(ws[i], LR, dw)
            bs[i] = o.apply_gradient_descent(bs[i], LR, db)
        return float(lvec.mean())

    return step, o.num_threads()


def cpu_baseline(max_seconds=15.0, max_steps=150):
    step, cores = cpu_step_fn()
    step()  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < max_seconds):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": BATCH * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d full training steps (batch %d, 3x1024 MLP, fwd+bwd+SGD) of the CPU "
                      "oracle port, OpenMP over %d host threads; %.2f s" % (n, BATCH, cores, dt)}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    step, cores = cpu_step_fn()
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))  # bounded: each step is a full 68.7 GFLOP CPU pass
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = BATCH * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mlp-3x1024 batch 4096 fp32 fwd+bwd+sgd (BASELINE configs[1])",
                   "note": "CPU restatement of the reference's Eigen path (reference unbuildable "
                           "offline: needs bazel+protoc+Eigen); steps bounded to %d" % steps},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d full training steps, %d OpenMP threads" % (steps, cores)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


# =================================================================================== GPU arm
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


LENET_BATCH = 512


def build_lenet_graph(num_replicas, seed):
    """BASELINE configs[2] / SURVEY 8d C3: conv5x5x1x32 SAME -> relu -> pool2 -> conv5x5x32x64 SAME
    -> relu -> pool2 -> fc 3136x1024 + relu -> fc 1024x10 -> xent; batch 512, NHWC fp32."""
    from simple_tensorflow_b200 import ops as tf
    rng = np.random.RandomState(seed)
    B = LENET_BATCH
    x = rng.uniform(-1, 1, (B, 28, 28, 1)).astype(np.float32)
    labels = np.zeros((B, 10), np.float32)
    labels[np.arange(B), rng.randint(0, 10, B)] = 1.0
    shapes = dict(w1=(5, 5, 1, 32), w2=(5, 5, 32, 64), w3=(3136, 1024), w4=(1024, 10))
    rng = np.random.RandomState(4321)  # identical initial weights on every replica
    tf.reset_default_graph()
    V = {}
    for n, shp in shapes.items():
        fan_in = int(np.prod(shp[:-1]))
        V[n] = tf.Variable((rng.randn(*shp) / np.sqrt(fan_in)).astype(np.float32), name=n)
        V["b" + n[1]] = tf.Variable(np.full(shp[-1], 0.1, np.float32), name="b" + n[1])
    train_vars = list(V.values())

    def tower(inp, lab, tag):
        c1 = tf.relu(tf.bias_add(tf.conv2d(inp, V["w1"], [1, 1, 1, 1], "SAME"), V["b1"]))
        p1 = tf.max_pool(c1, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
        c2 = tf.relu(tf.bias_add(tf.conv2d(p1, V["w2"], [1, 1, 1, 1], "SAME"), V["b2"]))
        p2 = tf.max_pool(c2, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
        flat = tf.reshape(p2, [B, 3136])
        f1 = tf.relu(tf.bias_add(tf.matmul(flat, V["w3"]), V["b3"]))
        logits = tf.bias_add(tf.matmul(f1, V["w4"]), V["b4"])
        loss = tf.reduce_mean(tf.softmax_cross_entropy_with_logits(logits, lab), name=tag + "/loss")
        train = tf.GradientDescentOptimizer(LR).minimize(loss, train_vars, name=tag + "/train",
                                                         num_replicas=num_replicas, **bucket_kw())
        return loss, train

    x_res = tf.Variable(x, name="x_resident")
    l_res = tf.Variable(labels, name="labels_resident")
    res = tower(x_res.ref, l_res.ref, "resident")
    xp = tf.placeholder(tf.float32, [B, 28, 28, 1], "x")
    lp = tf.placeholder(tf.float32, [B, 10], "labels")
    fed = tower(xp, lp, "fed")
    return tf, dict(x=x, labels=labels, xp=xp, lp=lp, resident=res, fed=fed)


def build_graph(num_replicas, seed):
    from simple_tensorflow_b200 import ops as tf
    x, labels, ws, bs = synthetic(seed)
    tf.reset_default_graph()
    Ws = [tf.Variable(w, name="W%d" % i) for i, w in enumerate(ws)]
    Bs = [tf.Variable(b, name="b%d" % i) for i, b in enumerate(bs)]

    def tower(inp, lab, tag):
        h = inp
        for i in range(LAYERS):
            h = tf.bias_add(tf.matmul(h, Ws[i], name="%s/fc%d" % (tag, i)), Bs[i])
            if i < LAYERS - 1:
                h = tf.relu(h)
        loss = tf.reduce_mean(tf.softmax_cross_entropy_with_logits(h, lab), name=tag + "/loss")
        train = tf.GradientDescentOptimizer(LR).minimize(loss, Ws + Bs, name=tag + "/train",
                                                         num_replicas=num_replicas, **bucket_kw())
        return loss, train

    # (a) inputs resident in HBM: device variables, assigned once
    x_res = tf.Variable(x, name="x_resident")
    l_res = tf.Variable(labels, name="labels_resident")
    # the towers must not train the data variables
    res = tower(x_res.ref, l_res.ref, "resident")
    # (b) host-fed placeholders
    xp = tf.placeholder(tf.float32, [BATCH, WIDTH], "x")
    lp = tf.placeholder(tf.float32, [BATCH, WIDTH], "labels")
    fed = tower(xp, lp, "fed")
    return tf, dict(x=x, labels=labels, xp=xp, lp=lp, resident=res, fed=fed)


def run_b200(args):
    # Libraries we load (NCCL's version banner) write to fd 1; the contract is ONE JSON line on
    # stdout, so everything else goes to stderr and the line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from simple_tensorflow_b200 import _lib, client
    L = _lib.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if L.b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libb200tf has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from simple_tensorflow_b200 import replica
        comm = replica.init_nccl_comm(L, rank, world, local_rank)

    lenet = args.workload == "lenet"
    batch = LENET_BATCH if lenet else BATCH
    tf, G = (build_lenet_graph if lenet else build_graph)(world, seed=1234 + rank)
    sess = client.Session(tf.get_default_graph(), gpu=local_rank, collective_comm=comm,
                          num_replicas=world)
    sess.run(tf.global_variables_initializer())
    stream = sess.stream()
    ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.b200_event_create(ctypes.byref(ev0)))
    _lib.check(L.b200_event_create(ctypes.byref(ev1)))
    hx = client.HostTensor.from_numpy(G["x"])        # pinned host buffers, reused every step
    hl = client.HostTensor.from_numpy(G["labels"])

    def barrier():
        _lib.check(L.b200_stream_synchronize(stream))
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def timed(fetches, feed, steps, prefetch=None):
        """-> (device ms between events on the session stream, launches, last loss).

        prefetch: list of (HostTensor x, HostTensor labels) buffer pairs used round-robin; the
        inputs of step i+1 are staged (Session.stage: H2D on the copy stream) before step i is
        run, so every step's host->device copy is inside the timed region but overlaps the
        previous step's kernels."""
        launches, loss, enq = 0, None, 0
        barrier()
        _lib.check(L.b200_event_record(ev0, stream))
        staged = None
        if prefetch:
            staged = tuple(sess.stage(t) for t in prefetch[0])
        for i in range(steps):
            if prefetch:
                nxt = (tuple(sess.stage(t) for t in prefetch[(i + 1) % len(prefetch)])
                       if i + 1 < steps else None)
                loss = sess.run(fetches, {G["xp"]: staged[0], G["lp"]: staged[1]})[0]
                staged = nxt
            else:
                loss = sess.run(fetches, feed)[0]
            st = sess.last_run_stats()
            launches += st["kernels_launched"]
            enq += st["host_enqueue_us"]
        timed.host_enqueue_us = enq / max(steps, 1)
        _lib.check(L.b200_event_record(ev1, stream))
        barrier()
        ms = ctypes.c_float()
        _lib.check(L.b200_event_elapsed_ms(ev0, ev1, ctypes.byref(ms)))
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms.value], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
            return float(t.item()), launches, float(loss)
        return ms.value, launches, float(loss)

    res_fetch = list(G["resident"])
    fed_fetch = list(G["fed"])
    feed = {G["xp"]: hx, G["lp"]: hl}
    # second pinned buffer pair: the input pipeline fills one while the other is in flight
    buffers = [(hx, hl), (client.HostTensor.from_numpy(G["x"]), client.HostTensor.from_numpy(G["labels"]))]
    warm = max(args.warmup, 3)
    for _ in range(warm):
        sess.run(res_fetch)
        sess.run(fed_fetch, feed)
    timed(fed_fetch, None, warm, prefetch=buffers)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, launches, loss_res = timed(res_fetch, None, args.steps)
    host_enqueue_us = timed.host_enqueue_us
    ms_sync, _, _ = timed(fed_fetch, feed, args.steps)  # feed pinned buffers, copy inside Run()
    h2d = sess.last_run_stats()["h2d_bytes"]
    ms_e2e, _, loss_e2e = timed(fed_fetch, None, args.steps, prefetch=buffers)
    clocks = sampler.summary() if rank == 0 else None
    assert sum(client.HostTensor.numpy(t).nbytes for t in buffers[0]) == h2d
    d2h = sess.last_run_stats()["d2h_bytes"]

    # ---- roofline pass: per-launch device time of the tcgen05 GEMM, events on the same stream
    _lib.check(L.b200_profile_begin())
    prof_steps = min(args.steps, 20)
    for _ in range(prof_steps):
        sess.run(res_fetch)
    gemm_ms, gemm_n, gemm_fl = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_double()
    _lib.check(L.b200_profile_end(ctypes.byref(gemm_ms), ctypes.byref(gemm_n), ctypes.byref(gemm_fl)))
    # what an event pair with nothing between costs on this stream (reported next to the raw
    # per-launch time: the bracket itself, not the kernel, explains why the kernel's share of the
    # step looks larger here than in the ncu launch list)
    pair = []
    for _ in range(50):
        _lib.check(L.b200_event_record(ev0, stream))
        _lib.check(L.b200_event_record(ev1, stream))
        _lib.check(L.b200_stream_synchronize(stream))
        ms_pair = ctypes.c_float()
        _lib.check(L.b200_event_elapsed_ms(ev0, ev1, ctypes.byref(ms_pair)))
        pair.append(ms_pair.value * 1e3)
    event_pair_us = statistics.median(pair)

    if rank != 0:
        sess.close()
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_sustained = peaks.get("bf16_tflops_sustained")
    if bf16_sustained:
        peak, peak_src = 0.5 * bf16_sustained, ("0.5 x measured sustained bf16 cuBLAS (%.1f TF): "
                                                "no measured TF32 figure in MEASURED_PEAKS.json"
                                                % bf16_sustained)
    else:
        peak, peak_src = 0.5 * 1400.0, "0.5 x fallback sustained bf16 1.4 PF (B200_PROFILING.md)"
    achieved = (gemm_fl.value / 1e12) / (gemm_ms.value / 1e3) if gemm_ms.value > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")))[
            "dram_bytes_per_launch"]
    except Exception:
        pass
    base = cpu_baseline() if world == 1 and not args.no_cpu_baseline and not lenet else None
    n_samples = batch * world * args.steps
    working_set_mb = (2 * BATCH * WIDTH * 4 + LAYERS * WIDTH * WIDTH * 4 * 2 +
                      8 * BATCH * WIDTH * 4) / 1e6
    line = {
        "metric": METRIC, "value": n_samples / (ms_res / 1e3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 graph; GEMMs on TF32 tensor cores, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": ("lenet-5 batch 512/replica NHWC fp32 fwd+bwd+sgd (BASELINE "
                                "configs[2]); Session.Run([loss, train_op])") if lenet else
                               ("mlp-3x1024 batch 4096/replica fp32 fwd+bwd+sgd "
                                "(BASELINE configs[1]); Session.Run([loss, train_op])"),
                   "global_batch": batch * world, "parallelism": "dp%d" % world,
                   "l2": ("no flush: per-step working set > 126 MB L2 (conv2 patch matrix alone is "
                          "321 MB)") if lenet else
                         ("no flush: per-step working set ~%.0f MB > 126 MB L2" % working_set_mb),
                   "loss_resident": loss_res, "loss_e2e": loss_e2e,
                   "host_enqueue_us_per_step": host_enqueue_us},
        "clocks": clocks,
        "e2e": {"value": n_samples / (ms_e2e / 1e3), "unit": "samples/s",
                "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "how": ("Session.stage() of step i+1's pinned inputs (copy stream) before "
                        "Session.run of step i; loss fetched to the host every step"),
                "unpipelined_value": n_samples / (ms_sync / 1e3),
                "unpipelined_ms_per_step": ms_sync / args.steps},
        "gpu_launches": launches,
        "roofline": {"kernel": "gemm_tcgen05_kernel (kind::tf32)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_source": peak_src, "launches_timed": int(gemm_n.value),
                     "us_per_launch": 1e3 * gemm_ms.value / max(1, gemm_n.value),
                     "empty_event_pair_us": event_pair_us,
                     "achieved_net_of_event_pair": ((gemm_fl.value / 1e12) /
                                                    max(1e-9, gemm_ms.value / 1e3 -
                                                        gemm_n.value * event_pair_us / 1e6)),
                     "flops_per_launch": (gemm_fl.value / max(1, gemm_n.value)),
                     "gemm_share_of_step": (gemm_ms.value / prof_steps) / (ms_res / args.steps)},
        "cpu_baseline": base,
    }
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    sess.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="mlp", choices=["mlp", "lenet"],
                    help="mlp = BASELINE configs[1] (default, the metric's config); lenet = configs[2]")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
