"""Helpers shared by the golden-vector tests (CPU oracle and GPU parity)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def iota(shape):
    """'incrementing numbers from 1' in row-major order (conv_ops_test.py:216-219)."""
    n = int(np.prod(shape, dtype=np.int64))
    return np.arange(1, n + 1, dtype=np.float32).reshape(shape)


def conv_cases(kind):
    return [r for r in load("conv_ops.json") if r["kind"] == kind]


def pool_cases(kind):
    out = []
    for r in load("pooling_ops.json"):
        if r["kind"] != kind:
            continue
        if kind == "max_pool" and (r["ksize"][3] != 1 or r["ksize"][0] != 1):
            continue  # depth-wise pooling is outside the hot path (pooling_ops_common.cc:54-57)
        out.append(r)
    return out


def case_id(r):
    return "%s@%s" % (r["test"], r["source"].rsplit(":", 1)[1])
