import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_device_count():
    try:
        from skeres_b200._lib import lib
        return lib.sk_device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def sk():
    """The product API; GPU tests fail (not skip) if the extension cannot run."""
    from skeres_b200 import api
    assert api.lib.sk_device_count() > 0, "no CUDA device: -m gpu tests must run on the GPU box"
    return api


def pytest_collection_modifyitems(config, items):
    # Without a GPU, `-m gpu` tests are skipped rather than failed when somebody runs the whole suite here.
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
