import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
    if p not in sys.path:
        sys.path.insert(0, p)

warnings.filterwarnings("ignore", category=UserWarning)
warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on a B200")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_yolo():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "yolo_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_rekt():
    import torch

    return torch.load(os.path.join(ROOT, "tests", "golden", "rektnet_golden.pt"), weights_only=False)["rektnet"]


@pytest.fixture(scope="session")
def cfg_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("cfgs"))
