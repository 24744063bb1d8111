import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/vcs_oracle.c) -- the checker, never the product."""
    from oracle import oracle
    oracle.build()
    return oracle
