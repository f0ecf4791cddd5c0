import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def zkb():
    import eigen_zeth_b200 as z
    return z


@pytest.fixture(scope="session")
def ctx(zkb):
    """one context for the whole GPU session; creating it fails loudly without an sm_100 GPU."""
    c = zkb.Context(0)
    yield c
    c.close()
