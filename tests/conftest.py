import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


@pytest.fixture(scope="session")
def evaluator():
    """The product path: C ABI over CUDA.  Fails loudly without a GPU."""
    from noisyquantumsimulator_b200.engine import Evaluator
    ev = Evaluator(0)
    yield ev
    ev.close()
