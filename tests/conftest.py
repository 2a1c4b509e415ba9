"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only builder container; `-m gpu` on a B200."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_oracle():
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], check=True,
                   stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference (fp32) -- the parity checker."""
    from ilqgames_b200 import _abi as abi
    _build_oracle()
    lib = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so"))
    lib.verify_struct_sizes()
    return lib


@pytest.fixture(scope="session")
def oracle64():
    """Same restatement with real = double: an accuracy yardstick for tolerances."""
    from ilqgames_b200 import _abi as abi
    _build_oracle()
    lib = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle64.so"))
    lib.verify_struct_sizes()
    return lib


@pytest.fixture(scope="session")
def product():
    """The sm_100a CUDA library through the C ABI.  No fallback: missing .so = failure."""
    from ilqgames_b200 import _abi as abi
    return abi.product_library()
