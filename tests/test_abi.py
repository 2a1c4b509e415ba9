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


def test_time_gated_records_fail_loudly_on_the_cuda_library(product_lib, oracle):
    """FinalTimeCost records (ilqg_cost_desc::active_from != 0) exist in the oracle only so far: the
    CUDA library must refuse them rather than ignore the gate."""
    desc, _ = problems.two_player_collision()
    assert sum(1 for c in range(desc.num_costs) if desc.costs[c].active_from != 0.0) == 4
    p = problems.two_player_collision_params()
    h = C.c_void_p()
    assert product_lib.lib.ilqg_create(C.byref(desc), C.byref(p), 1, 0, C.byref(h)) == -2  # ILQG_ERR_UNSUPPORTED
    assert oracle.lib.ilqg_create(C.byref(desc), C.byref(p), 1, 0, C.byref(h)) == 0
    assert oracle.lib.ilqg_destroy(h) == 0
    # likewise SinglePlayerCar5D / SignedDistanceCost (oracle-only kinds so far)
    desc5, _ = problems.two_player_collision_avoidance_reachability()
    p5 = problems.two_player_collision_avoidance_reachability_params()
    assert product_lib.lib.ilqg_create(C.byref(desc5), C.byref(p5), 1, 0, C.byref(h)) == -2
    assert oracle.lib.ilqg_create(C.byref(desc5), C.byref(p5), 1, 0, C.byref(h)) == 0
    assert oracle.lib.ilqg_destroy(h) == 0
    descpm, _ = problems.modified_air_3d()                       # SinglePlayerPointMass2D
    assert product_lib.lib.ilqg_create(C.byref(descpm), C.byref(p5), 1, 0, C.byref(h)) == -2
    desc2p, _ = problems.two_player_reachability()               # TwoPlayerUnicycle4D
    assert product_lib.lib.ilqg_create(C.byref(desc2p), C.byref(p5), 1, 0, C.byref(h)) == -2
    desc1, _ = problems.one_player_reachability()                # SinglePlayerDubinsCar
    assert product_lib.lib.ilqg_create(C.byref(desc1), C.byref(p5), 1, 0, C.byref(h)) == -2
    # ... and ExtremeValueCost groups; a group member that is a constraint or gated is invalid
    desc3, _ = problems.three_player_collision_avoidance_reachability()
    assert sorted({desc3.costs[c].group for c in range(desc3.num_costs)}) == [0, 1, 2, 3]
    assert oracle.lib.ilqg_create(C.byref(desc3), C.byref(p5), 1, 0, C.byref(h)) == 0
    assert oracle.lib.ilqg_destroy(h) == 0
    for c in range(desc3.num_subsystems):
        desc3.subsystems[c].kind = abi.DYN_UNICYCLE4D      # leave only the groups for the CUDA parser to object to
    desc3.xdim = 12
    for c in range(desc3.num_subsystems):
        desc3.subsystems[c].x_offset = 4 * c
    for c in range(desc3.num_costs):
        if desc3.costs[c].kind == abi.COST_SIGNED_DISTANCE:
            desc3.costs[c].kind = abi.COST_PROXIMITY
            for q in range(4):
                desc3.costs[c].dim[q] = desc3.costs[c].dim[q] // 5 * 4 + desc3.costs[c].dim[q] % 5
    assert product_lib.lib.ilqg_create(C.byref(desc3), C.byref(p5), 1, 0, C.byref(h)) == -2
    for c in range(desc3.num_costs):
        desc3.costs[c].group = 0
    rc = product_lib.lib.ilqg_create(C.byref(desc3), C.byref(p5), 1, 0, C.byref(h))
    assert rc in (-4, 0)   # parses: "no device" on the CPU-only builder, a handle on a GPU box
    if rc == 0:
        assert product_lib.lib.ilqg_destroy(h) == 0
    # a gate on a constraint record is not a thing in the reference (FinalTimeCost wraps Costs)
    bad, _ = problems.three_player_intersection()
    for c in range(bad.num_costs):
        if bad.costs[c].kind == abi.CONSTRAINT_PROXIMITY:
            bad.costs[c].active_from = 1.0
    assert oracle.lib.ilqg_create(C.byref(bad), C.byref(p), 1, 0, C.byref(h)) == -1


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
