import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


ORACLE_LIB = ROOT / "oracle" / "libmor_oracle.so"


def _ensure_built():
    """Build the oracle / synth / product libraries if they are missing (CPU-only; nvcc cross-compiles)."""
    from dynamicslamtool_b200.binding import PRODUCT_LIB, SYNTH_LIB
    if not ORACLE_LIB.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle")], stdout=subprocess.DEVNULL)
    if not SYNTH_LIB.exists() or not PRODUCT_LIB.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "dynamicslamtool_b200" / "csrc")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def built():
    _ensure_built()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    """The CPU oracle binding (test infrastructure only)."""
    from dynamicslamtool_b200.binding import MorBinding
    return MorBinding(C.CDLL(str(ORACLE_LIB)), "oracle_")


@pytest.fixture(scope="session")
def product(built):
    from dynamicslamtool_b200.binding import load_product
    return load_product()


@pytest.fixture(scope="session")
def cfg_dir():
    return ROOT / "config"
