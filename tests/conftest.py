import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def synthetic_sd():
    """The deterministic 75 M-parameter synthetic state_dict (oracle/weights.py)."""
    from oracle import weights
    return weights.make_state_dict(seed=0)


@pytest.fixture(scope="session")
def scale_table():
    from oracle import weights
    return weights.scale_table()
