import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hyper-vla_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def params_p1():
    from hvla import params as P
    return P.init_params(2025, "P1")


@pytest.fixture(scope="session")
def params_p0():
    from hvla import params as P
    return P.init_params(2025, "P0")


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    return {n[:-4]: np.load(os.path.join(d, n)) for n in sorted(os.listdir(d)) if n.endswith(".npz")}
