import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # tests may import the oracle (the checker); the product never does


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _cuda_device_present() -> bool:
    """true when the engine can create a context (it has no CPU fallback: QTB_ERR_NO_DEVICE otherwise)"""
    try:
        import quantit_b200 as qb

        qb.load_library()
        qb.default_context()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing them; the loud failure
    # of the product itself without a device is asserted by tests/test_capi_symbols.py (not gpu-marked)
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _cuda_device_present():
        skip = pytest.mark.skip(reason="no CUDA device: the qtb engine has no CPU fallback")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    import quantit_b200 as qb

    qb.load_library()
    return qb
