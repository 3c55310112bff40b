"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs the reference's golden vectors, host logic, ABI.
`-m gpu` runs on a B200: parity of the CUDA path (through the C ABI) against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The shared libraries are build products (git-ignored): a fresh checkout builds them once
    # (nvcc cross-compiles sm_100a without a GPU); on the GPU box the snapshot carries them.
    lib_dir = os.path.join(ROOT, "simple_tensorflow_b200", "lib")
    if not (os.path.exists(os.path.join(lib_dir, "libb200tf.so")) and
            os.path.exists(os.path.join(lib_dir, "libb200tf_framework.so"))):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    import oracle_bind
    oracle_bind.lib()
    return oracle_bind


@pytest.fixture
def rng(request):
    """A RandomState seeded from the test's own node id: inputs do not depend on which other
    tests ran before (selection / ordering / -x)."""
    import zlib
    return np.random.RandomState(zlib.crc32(request.node.nodeid.encode()) & 0x7FFFFFFF)
