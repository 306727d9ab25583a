import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import wc_oracle

    wc_oracle.build()
    return wc_oracle


@pytest.fixture(scope="session")
def c1_window():
    from wildcat_slam_b200 import synthetic

    return synthetic.make_window("C1")


@pytest.fixture(scope="session")
def c2_window():
    from wildcat_slam_b200 import synthetic

    return synthetic.make_window("C2")
