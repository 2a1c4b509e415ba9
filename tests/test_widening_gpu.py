"""GPU tests of the widening -- kept in a file that sorts after every other GPU test
file on purpose: the K_bwd instance for n = 18 below has not run on a device yet, and a kernel that
faulted would leave the process's CUDA context unusable for whatever came after it.

The instance (the warp-per-game kernel; 18 is not a multiple of 4, which the half-warp kernel
needs) was added after round 1's GPU minutes were spent: the CPU side (oracle bit for bit against the
reference, tests/test_ref_pins.py; the example source compiling unchanged, tests/test_host_api.py)
is verified, the device side runs for the first time in the driver's own GPU pass.  Not strict: a
pass is reported as XPASS."""
import pytest

import tests.test_gpu_parity as parity

pytestmark = pytest.mark.gpu

FIRST_DEVICE_RUN = pytest.mark.xfail(strict=False, reason="n = 18 K_bwd instance not yet run on a device")


@FIRST_DEVICE_RUN
def test_overtaking_stage_parity(product, oracle, oracle64):
    # one iteration: after it these games sit at the merit floor, where the number of backtracking
    # steps is decided by rounding (the oracle's own fp32 and fp64 builds take 18 and 48)
    parity.test_stage_parity(product, oracle, oracle64, "three_player_overtaking", iterations=1)


@FIRST_DEVICE_RUN
def test_overtaking_against_reference_fixture(product, oracle64):
    parity.test_against_reference_fixture(product, oracle64, "three_player_overtaking")
