import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def load_weights(ckpt):
    return torch.load(os.path.join(GOLD, ckpt), map_location="cpu", weights_only=False)["model"]


@pytest.fixture(scope="session")
def ops_dtu():
    return load_golden("ops_dtu.pt")


@pytest.fixture(scope="session")
def sd_dtu():
    return load_weights("model_dtu_v2.ckpt")
