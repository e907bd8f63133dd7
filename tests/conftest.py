import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import avsum_b200  # noqa: E402,F401  (registers the package under its importable name)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def native_lib():
    from avsum_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    return _cabi.lib()


@pytest.fixture(scope="session")
def cuda_ready(native_lib):
    import torch
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    assert native_lib.avs_device_ok() == 1, "libavsum_b200.so sees no sm_100 device"
    return True
