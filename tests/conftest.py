"""pytest configuration: the ``gpu`` marker and shared helpers."""
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name: str):
    """Load ``tests/golden/<name>.npz`` as a dict of torch tensors."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as data:
        return {key: torch.from_numpy(data[key]) for key in data.files}


def sub_dict(golden: dict, prefix: str) -> dict:
    """Entries of ``golden`` under ``prefix.`` with the prefix stripped."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in golden.items() if k.startswith(prefix + ".")}


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
