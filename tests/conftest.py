import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


def _gpu_visible() -> bool:
    """True when a CUDA device can be opened -- asked through the product library itself (rydcz_create), so the
    answer is the one the tests would get.  If the library is not built, say True: the gpu tests must then fail
    loudly on the missing library rather than be skipped."""
    import ctypes
    try:
        from noisyquantumsimulator_b200 import _lib
        lib = _lib.load()
    except Exception:
        return True
    ctx = ctypes.c_void_p()
    if lib.rydcz_create(0, ctypes.byref(ctx)) != 0:
        return False
    lib.rydcz_destroy(ctx)
    return True


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a GPU skips the gpu-marked tests instead of failing at the
    first one (`-m gpu` on the B200 box runs them; the product path itself still raises without a GPU)."""
    if _gpu_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def evaluator():
    """The product path: C ABI over CUDA.  Fails loudly without a GPU."""
    from noisyquantumsimulator_b200.engine import Evaluator
    ev = Evaluator(0)
    yield ev
    ev.close()
