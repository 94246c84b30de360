import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oq():
    import oetqf_b200
    return oetqf_b200


@pytest.fixture(scope="session")
def gpu(oq):
    """Initialise the device once; the product has no CPU fallback, so this fails loudly without a GPU."""
    oq.init(0)
    return oq
