import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_ops():
    return dict(np.load(os.path.join(GOLDEN, "ops.npz")))


@pytest.fixture(scope="session")
def golden_g32():
    return dict(np.load(os.path.join(GOLDEN, "generator32.npz")))


@pytest.fixture(scope="session")
def golden_g128():
    return dict(np.load(os.path.join(GOLDEN, "generator128.npz")))


def max_abs(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return (a - b).abs().max().item()


def psnr_db(test, ref, peak=2.0):
    test = torch.as_tensor(test).double()
    ref = torch.as_tensor(ref).double()
    mse = ((test - ref) ** 2).mean().item()
    return float("inf") if mse == 0 else 10 * np.log10(peak * peak / mse)
