"""pytest configuration: `gpu` marker (needs a real B200) and shared fixtures.

`python -m pytest tests -m "not gpu"` runs on CPU (oracle vs golden vectors, host logic, C-ABI symbol check);
`python -m pytest tests -m gpu` runs the CUDA parity tests through the C ABI on a B200.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with `-m gpu`")


@pytest.fixture(scope="session")
def gold():
    return np.load(GOLDEN / "reference_outputs.npz")


@pytest.fixture(scope="session")
def gold_vitl():
    """Reference outputs at the ViT-L/14 widths / sequence lengths (tests/golden/make_golden_vitl.py)."""
    return np.load(GOLDEN / "reference_outputs_vitl.npz")


@pytest.fixture(scope="session")
def gold_full():
    """Reference outputs at FULL depth for BASELINE.json configs 1, 3, 4 (tests/golden/make_golden_full.py)."""
    return np.load(GOLDEN / "reference_outputs_full.npz")


@pytest.fixture(scope="session")
def meta():
    return json.loads((GOLDEN / "reference_meta.json").read_text())


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a `gpu` test was selected but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")
