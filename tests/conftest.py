"""pytest configuration: markers, paths, shared fixtures.

`-m "not gpu"` tests run in the build container (no GPU): the oracle against the
reference's golden vectors, the per-cell device formulas compiled for the host
(tests/host_emu), host logic, the C-ABI symbol table and the world_size-2 gloo
tests.  `-m gpu` tests are the parity tests proper and go through the C ABI.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(GOLDEN, "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_exec():
    return dict(np.load(os.path.join(GOLDEN, "ref_exec.npz")))


@pytest.fixture(scope="session")
def rxj_data():
    return dict(np.load(os.path.join(GOLDEN, "rxj1713_data.npz")))


@pytest.fixture(scope="session")
def lut_probe():
    return dict(np.load(os.path.join(GOLDEN, "pp_lut_probe.npz")))


@pytest.fixture(scope="session")
def ref_units():
    """Outputs of the reference's unit-bound functions (tests/golden/make_golden_units.py)."""
    return dict(np.load(os.path.join(GOLDEN, "ref_exec_units.npz")))
