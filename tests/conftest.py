import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cm():
    from mujocoposelearning_b200.mjcf import compile_mjcf
    return compile_mjcf()


@pytest.fixture(scope="session")
def model_struct(cm):
    from mujocoposelearning_b200.abi import pack_model
    return pack_model(cm)
