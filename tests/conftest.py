import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    """Number of CUDA devices the product library sees (0 if it is not built or there is no driver)."""
    try:
        from bumpcosmology_b200 import _lib
        return int(_lib.load().bump_device_count())
    except Exception:  # noqa: BLE001  (library missing: the gpu tests cannot run either)
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without CUDA: skip the gpu-marked tests instead of failing them (the product has no CPU
    path, so they cannot pass there).  On the B200 box nothing is skipped: `-m gpu` runs them all."""
    if not any("gpu" in item.keywords for item in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (bump_device_count() == 0)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
