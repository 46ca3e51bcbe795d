"""pytest configuration: `gpu` marker + shared fixtures.

CPU suite (`-m "not gpu"`): oracle port vs oracle/_ref vs committed golden vectors, host logic, C-ABI symbols.
GPU suite (`-m gpu`): the CUDA path, called through the C ABI, against the oracle on the same inputs.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")  # sequential reductions => the two CPU checkers agree bitwise

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_runs():
    with open(os.path.join(GOLDEN, "runs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_fields():
    return np.load(os.path.join(GOLDEN, "fields.npz"))


@pytest.fixture(scope="session")
def golden_stages():
    return np.load(os.path.join(GOLDEN, "stages.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle import cfdo
    return cfdo.load("port")


@pytest.fixture(scope="session")
def ref():
    from oracle import cfdo
    if not cfdo.have_ref():
        pytest.skip("oracle/_ref not available (no /root/reference and no prebuilt .so)")
    return cfdo.load("ref")
