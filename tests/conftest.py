import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu() -> bool:
    try:
        import graphblas_b200 as g
        return g.lib.gb200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()


@pytest.fixture(scope="session")
def G():
    """The reference GraphBLAS library with the B200 shim interposed (switched off by default)."""
    import grbref
    if not grbref.available():
        pytest.skip("oracle/_ref/libgraphblas_ref.so not built")
    g = grbref.GraphBLAS.get(with_shim=True)
    g.use_gpu(False)
    return g
