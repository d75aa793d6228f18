"""pytest configuration: registers the `gpu` marker and puts the package on sys.path.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol table (no GPU needed).
`-m gpu`      : parity tests proper - CUDA path through the C-ABI vs oracle / golden.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "yolo-mslesseg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(GOLDEN_DIR / "golden_v1.json") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def demo_slices():
    return np.load(GOLDEN_DIR / "demo_slices.npz")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
