import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tb():
    return importlib.import_module("t-one_b200")


@pytest.fixture(scope="session")
def weights(tb):
    return tb.weights.init_weights(seed=0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {ms: np.load(os.path.join(d, f"step_{ms}ms.npz")) for ms in (300, 400)}
