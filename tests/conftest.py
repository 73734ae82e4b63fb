import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "reference_golden.npz", allow_pickle=False)


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The C checker (oracle/liboracle.so) is test infrastructure: build it if it is missing."""
    if not (ROOT / "oracle" / "liboracle.so").exists():
        from oracle import cpu
        cpu.build()
