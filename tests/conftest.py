import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def goldens():
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def goldens_heads2():
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens_heads2.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def goldens_xfusion4():
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens_xfusion4.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def goldens_mm():
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens_mm.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def goldens_unimodal():
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens_unimodal.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def goldens_losses():
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens_losses.pt")
    return torch.load(path, map_location="cpu", weights_only=False)
