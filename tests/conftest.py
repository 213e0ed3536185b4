import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def small_store():
    from oarfish_b200 import synth
    return synth.make_config("small", want_truth=True)


@pytest.fixture(scope="session")
def tiny_store():
    from oarfish_b200 import synth
    return synth.make_config("tiny", want_truth=True)
