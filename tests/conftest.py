import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from clearml_serving_b200 import native
        return native.device_count() > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not HAS_GPU:
            item.add_marker(skip_gpu)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def gpu_native():
    """Builds (if stale) and initialises libb200serve on cuda:0. GPU tests fail -- not skip --
    when the extension cannot be used on a GPU box."""
    from clearml_serving_b200 import build, native
    build.build()
    native.ensure_init(0)
    return native
