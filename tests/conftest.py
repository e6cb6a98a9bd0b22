import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "reference_sampler.npz")


@pytest.fixture(scope="session")
def lib():
    """Builds (if needed) and loads the C-ABI library."""
    from arcflow_b200 import build, _lib
    build.build()
    return _lib.load()
