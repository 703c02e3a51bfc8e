import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    # Speedy.set_bc() without sst_anomaly= warns that the reference's default anomaly file is not packaged (zero anomaly)
    config.addinivalue_line("filterwarnings", "ignore:pyspeedy_b200. the default SST anomaly file:RuntimeWarning")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def drv():
    """The CUDA library through its C ABI (ctypes); fails loudly if it was not built."""
    from pyspeedy_b200 import _driver

    return _driver
