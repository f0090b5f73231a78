import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def scorer():
    """One device handle for the whole GPU session (buffers are reused between problems)."""
    from cge_jl_b200.divergence import Scorer

    sc = Scorer(0)
    yield sc
    sc.close()
