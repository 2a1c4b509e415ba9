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


@pytest.mark.parametrize("name", parity.WIDENED)
def test_widened_example_stage_parity(product, oracle, oracle64, name):
    """Ten more examples of the reference on the device (round 2): SinglePlayerCar5D, SinglePlayerDubinsCar,
    SinglePlayerPointMass2D, TwoPlayerUnicycle4D; SignedDistanceCost, QuadraticDifferenceCost, FinalTimeCost
    gates, ExtremeValueCost groups.  One iteration, stage by stage, against the oracle (which reproduces
    the reference's own sources bit for bit on these problems, tests/test_ref_pins.py)."""
    # (the fixtures' perturbed initial states make several of these examples wild -- merits of 1e15 .. 1e37 -- so
    # only a few of the 8 games are well posed in fp32: TwoPlayerCollision keeps one)
    parity.test_stage_parity(product, oracle, oracle64, name, 100, iterations=1, min_wellposed=1)


@pytest.mark.parametrize("name", parity.WIDENED)
def test_widened_example_against_reference_fixture(product, oracle64, name):
    parity.test_against_reference_fixture(product, oracle64, name, min_compared=1)
