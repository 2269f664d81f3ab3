import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        from mpc_rs_b200 import _abi
        return _abi.lib().mpcb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def gpu_required():
    if not _has_gpu():
        pytest.fail("no CUDA device / libmpc_b200.so: the GPU tests cannot run (there is no CPU fallback)")
    return True
