import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "wan2.1-quantization_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
try:
    importlib.import_module("omegaconf")
except ImportError:                      # config container only; see compat/omegaconf/__init__.py
    sys.path.insert(0, os.path.join(PKG, "compat"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
