import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_drand48():
    return dict(np.load(os.path.join(GOLDEN, "laplace_drand48_n3000_p4.npz")))


@pytest.fixture(scope="session")
def golden_two_scale():
    return dict(np.load(os.path.join(GOLDEN, "laplace_two_scale_n4000_p6.npz")))


@pytest.fixture(scope="session")
def checksums():
    import json
    return json.load(open(os.path.join(GOLDEN, "checksums.json")))
