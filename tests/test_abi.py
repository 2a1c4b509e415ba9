"""CPU-side checks of the C-ABI boundary: the product library loads, exports every symbol
include/ilqg.h declares, agrees with the ctypes mirror on struct sizes, validates descriptors,
and refuses to compute without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ilqgames_b200 import _abi as abi
from ilqgames_b200 import build as b
from ilqgames_b200 import problems

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    b.build()
    lib = abi.Library(abi.PRODUCT_LIB)
    return lib


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "ilqg.h")).read()
    return sorted(set(re.findall(r"\b(ilqg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_match_binding_table():
    assert _declared_symbols() == sorted(abi.ABI_SYMBOLS)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_library_exports_every_declared_symbol(which, product_lib, oracle):
    lib = product_lib if which == "product" else oracle
    for sym in _declared_symbols():
        assert hasattr(lib.lib, sym), f"{lib.path} does not export {sym}"
    lib.verify_struct_sizes()


def test_product_has_no_cpu_fallback(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    desc, _ = problems.three_player_intersection()
    h = C.c_void_p()
    rc = product_lib.lib.ilqg_create(C.byref(desc), C.byref(abi.SolverParams.defaults()), 4, 0, C.byref(h))
    assert rc == -4  # ILQG_ERR_NO_DEVICE
    assert b"no CUDA device" in product_lib.lib.ilqg_strerror(rc)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(abi.IlqgError, match="no CPU fallback"):
        abi.Library(str(tmp_path / "libilqg_b200.so"))


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_invalid_descriptors_are_rejected(which, product_lib, oracle):
    lib = product_lib if which == "product" else oracle
    p = abi.SolverParams.defaults()
    h = C.c_void_p()
    desc, _ = problems.three_player_intersection()
    assert lib.lib.ilqg_create(None, C.byref(p), 1, 0, C.byref(h)) == -1
    assert lib.lib.ilqg_create(C.byref(desc), C.byref(p), 0, 0, C.byref(h)) == -1
    bad = abi.ProblemDesc.from_buffer_copy(desc)
    bad.num_time_steps = 1
    assert lib.lib.ilqg_create(C.byref(bad), C.byref(p), 1, 0, C.byref(h)) == -1
    bad = abi.ProblemDesc.from_buffer_copy(desc)
    bad.costs[0].player = 7
    assert lib.lib.ilqg_create(C.byref(bad), C.byref(p), 1, 0, C.byref(h)) == -1
    if which == "oracle":
        ol = abi.SolverParams.defaults(open_loop=1)  # LQOpenLoopSolver (ilq_solver.h:76-81)
        assert lib.lib.ilqg_create(C.byref(desc), C.byref(ol), 1, 0, C.byref(h)) == 0
        assert lib.lib.ilqg_destroy(h) == 0


def test_widened_descriptors_parse_on_the_cuda_library(product_lib, oracle):
    """Row f4: FinalTimeCost gates, ExtremeValueCost groups, SinglePlayerCar5D / DubinsCar /
    PointMass2D, TwoPlayerUnicycle4D, SignedDistanceCost and QuadraticDifferenceCost were oracle-only
    in round 1 (the CUDA library answered ILQG_ERR_UNSUPPORTED).  The CUDA library's parser now takes
    them: "no device" (-4) on the CPU-only builder, a handle on a GPU box -- never -2 / -1."""
    h = C.c_void_p()
    p = abi.SolverParams.defaults()

    def accepted(lib, desc, params):
        rc = lib.lib.ilqg_create(C.byref(desc), C.byref(params), 1, 0, C.byref(h))
        if rc == 0:
            assert lib.lib.ilqg_destroy(h) == 0
        return rc

    desc, _ = problems.two_player_collision()           # FinalTimeCost gates
    assert sum(1 for c in range(desc.num_costs) if desc.costs[c].active_from != 0.0) == 4
    assert accepted(product_lib, desc, problems.two_player_collision_params()) in (0, -4)
    assert accepted(oracle, desc, problems.two_player_collision_params()) == 0
    p5 = problems.two_player_collision_avoidance_reachability_params()
    for build in (problems.two_player_collision_avoidance_reachability,    # Car5D, SignedDistanceCost
                  problems.modified_air_3d,                                  # SinglePlayerPointMass2D
                  problems.two_player_reachability,                          # TwoPlayerUnicycle4D
                  problems.one_player_reachability,                          # SinglePlayerDubinsCar
                  problems.dubins_origin,                                    # QuadraticDifferenceCost
                  problems.three_player_collision_avoidance_reachability):   # ExtremeValueCost groups
        d, _ = build()
        assert accepted(product_lib, d, p5) in (0, -4), build.__name__
        assert accepted(oracle, d, p5) == 0, build.__name__
    desc3, _ = problems.three_player_collision_avoidance_reachability()
    assert sorted({desc3.costs[c].group for c in range(desc3.num_costs)}) == [0, 1, 2, 3]
    # a gate on a constraint record is not a thing in the reference (FinalTimeCost wraps Costs); a group
    # member that is gated is invalid as well -- both libraries say so
    bad, _ = problems.three_player_intersection()
    for c in range(bad.num_costs):
        if bad.costs[c].kind == abi.CONSTRAINT_PROXIMITY:
            bad.costs[c].active_from = 1.0
    assert oracle.lib.ilqg_create(C.byref(bad), C.byref(p), 1, 0, C.byref(h)) == -1
    assert product_lib.lib.ilqg_create(C.byref(bad), C.byref(p), 1, 0, C.byref(h)) == -1
    for c in range(desc3.num_costs):
        if desc3.costs[c].group > 0:
            desc3.costs[c].active_from = 1.0
            break
    assert oracle.lib.ilqg_create(C.byref(desc3), C.byref(p5), 1, 0, C.byref(h)) == -1
    assert product_lib.lib.ilqg_create(C.byref(desc3), C.byref(p5), 1, 0, C.byref(h)) == -1
    # polyline ranges that overlap / overflow the segment table are refused
    over, _ = problems.three_player_intersection()
    for q in range(over.num_polylines + 1):
        over.polyline_start[q] = 0 if q == 0 else 100
    assert product_lib.lib.ilqg_create(C.byref(over), C.byref(p), 1, 0, C.byref(h)) == -1
    neg, _ = problems.three_player_intersection()
    neg.costs[0].arg = -2
    assert product_lib.lib.ilqg_create(C.byref(neg), C.byref(p), 1, 0, C.byref(h)) == -1


def test_descriptor_shapes_of_the_three_configs(oracle):
    # SURVEY.md section 8 shape table (reference-true shapes)
    for build, N, n, M, ncon, ncost in ((problems.three_player_intersection, 3, 16, 6, 6, 18),
                                        (problems.roundabout_merging, 4, 24, 8, 0, 44),
                                        (problems.air_3d, 2, 3, 2, 4, 8)):
        desc, x0 = build()
        h = abi.Handle(oracle, desc, abi.SolverParams.defaults(), 1)
        assert (h.N, h.n, h.M, h.T) == (N, n, M, 100)
        assert h.layout.num_constraints == ncon
        assert desc.num_costs == ncost
        assert x0.shape == (n,)
        h.close()
