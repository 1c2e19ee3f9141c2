import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SITE = os.path.join(ROOT, "temporal-pointcloud-upsampling-gan_b200")
for p in (ROOT, SITE, os.path.join(ROOT, "tools"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests never silently pass on a box without a GPU: they are skipped with a
    reason unless selected with -m gpu, in which case a missing device is a failure."""
    import torch

    if torch.cuda.is_available():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    for item in items:
        if "gpu" in item.keywords and not selected_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device in this container"))


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle

    _oracle.build()
    return _oracle
