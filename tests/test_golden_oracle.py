"""The committed golden fixtures are what the current oracle produces (guards against silent
drift of the checker; regenerate with tests/golden/make_golden.py when the oracle changes on
purpose)."""
import os

import numpy as np
import pytest

from tests.golden import make_golden


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(oracle, oracle64, name):
    g = np.load(os.path.join(os.path.dirname(make_golden.__file__), f"{name}.npz"))
    out = make_golden.run_case(oracle, name, oracle64)
    assert sorted(out) == sorted(g.files)
    for k in g.files:
        np.testing.assert_array_equal(out[k], g[k], err_msg=k)
