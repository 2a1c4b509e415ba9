"""GPU tests of the widening: examples of the reference beyond the three headline ones,
on the device.  ThreePlayerOvertaking (3 x Car6D, n = 18) first ran on a B200 in the driver's round-1
GPU pass (both tests passed; the non-strict xfail marks they carried then are gone).  Since round 2
these shapes run the tensor-core sweep over compact records (n = 18 padded to 24)."""
import pytest

import tests.test_gpu_parity as parity

pytestmark = pytest.mark.gpu



def test_overtaking_stage_parity(product, oracle, oracle64):
    # one iteration: after it these games sit at the merit floor, where the number of backtracking
    # steps is decided by rounding (the oracle's own fp32 and fp64 builds take 18 and 48)
    parity.test_stage_parity(product, oracle, oracle64, "three_player_overtaking", 100, iterations=1)


def test_overtaking_against_reference_fixture(product, oracle64):
    parity.test_against_reference_fixture(product, oracle64, "three_player_overtaking")
