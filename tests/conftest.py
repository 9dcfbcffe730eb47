import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from yt_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _has_gpu():
        pytest.fail("GPU test selected but libytb200.so / a CUDA device is not available")
    from yt_b200 import device
    return device.get_context()


@pytest.fixture(autouse=True)
def _release_device_memory(request):
    """The full-size tests hold 4-30 GB of HBM each through objects that sit in reference cycles (sampler <-> frame <-> the
    tensors it keeps alive); Python frees those only when the cyclic collector happens to run, so a file of such tests run on
    its own could exhaust 180 GB (seen under compute-sanitizer, where nothing else allocates in between).  Collect after every
    GPU test and hand the cached blocks back."""
    yield
    if request.node.get_closest_marker("gpu") is not None:
        import gc
        gc.collect()
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        except Exception:   # noqa: BLE001
            pass
