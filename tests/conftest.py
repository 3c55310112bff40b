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


@pytest.fixture(scope="session")
def oracle():
    import oracle_bind
    oracle_bind.lib()
    return oracle_bind


@pytest.fixture(scope="session")
def rng():
    return np.random.RandomState(1234)
