"""The CPU oracle against outputs of THE REFERENCE ITSELF.

tests/golden/ref_*.npz were produced by tests/golden/make_ref_golden.py from the reference's own
sources (/root/reference/src/*.cpp compiled unmodified against the Eigen/glog/gflags stand-ins in
oracle/ref_shim; see oracle/ref_shim/Eigen/Dense for what that pins).  They hold what
ILQSolver::Solve and AugmentedLagrangianSolver::Solve return on the reference's own example
Problems: every logged iterate, the final strategies, total costs, success flags, iterate counts,
multipliers.  The oracle has to reproduce them BIT FOR BIT -- trajectories, feedback gains,
statuses and counters alike; only the per-player total costs get one ulp of slack (their summation
order follows an unordered_multimap in the reference).

The GPU parity tests (tests/test_gpu_parity.py::test_against_reference_fixture) hold the CUDA path
to the same fixtures within the fp32 tolerances stated there."""
import os

import numpy as np
import pytest

from ilqgames_b200 import _abi as abi, al, problems

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

CASES = {
    "three_player_intersection": (problems.three_player_intersection,
                                  problems.three_player_intersection_params),
    "roundabout_merging": (problems.roundabout_merging, problems.roundabout_params),
    "air_3d": (problems.air_3d, problems.air_3d_params),
    "three_player_overtaking": (problems.three_player_overtaking, problems.three_player_overtaking_params),
    "two_player_collision": (problems.two_player_collision, problems.two_player_collision_params),
    "two_player_collision_avoidance_reachability": (problems.two_player_collision_avoidance_reachability,
                                                    problems.two_player_collision_avoidance_reachability_params),
    "three_player_collision_avoidance_reachability": (problems.three_player_collision_avoidance_reachability,
                                                      problems.three_player_collision_avoidance_reachability_params),
    "one_player_reachability": (problems.one_player_reachability, problems.one_player_reachability_params),
    "dubins_origin": (problems.dubins_origin, problems.dubins_origin_params),
    "two_player_reachability": (problems.two_player_reachability, problems.two_player_reachability_params),
    "modified_air_3d": (problems.modified_air_3d, problems.modified_air_3d_params),
}


def load(name):
    return np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))


def _from_fixture(name):
    """Examples without a hand-written builder in problems.py: the fixture carries the descriptor
    that include/ilqgames/**'s DescribeProblem emitted from the example's own source
    (tests/golden/make_ref_golden.py:source_descriptor) and the executable's solver parameters."""
    def build():
        g = load(name)
        return abi.ProblemDesc.from_buffer_copy(g["desc_bytes"].tobytes()), g["x0"][0].copy()

    def params(**overrides):
        alpha, tol, decrease = (float(v) for v in load(name)["solver_params"])
        base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=decrease,
                    initial_alpha_scaling=alpha, convergence_tolerance=tol)
        base.update(overrides)
        return abi.SolverParams.defaults(**base)
    return build, params


for _name in ("modified_three_player_intersection", "skeleton", "three_player_intersection_reachability"):
    CASES[_name] = _from_fixture(_name)


def sorted_rows(a):
    return a[np.lexsort(a.T[::-1])] if a.shape[0] else a


def ulp_close(a, b, ulps):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)))))


def test_descriptor_polylines_equal_the_references():
    """RoundaboutLaneCenter / DrawCircle as the reference computes them (cosf/sinf)."""
    import math
    g = np.load(os.path.join(GOLDEN, "ref_polylines.npz"))
    F = np.float32
    off, wedge = F(math.pi / 2 * 0.5), F(math.pi)
    for i, dist in enumerate((25.0, 10.0, 25.0, 10.0)):
        ang = F(float(off) + i * 2.0 * math.pi / 4.0)
        mine = np.asarray(problems.roundabout_lane_center(ang, F(ang + wedge), dist), np.float32)
        assert np.array_equal(mine, g[f"roundabout_lane_{i}"])
    mine = np.asarray(problems.draw_circle((0.0, 0.0), 5.0, 10), np.float32)
    assert np.array_equal(mine, g["air_3d_circle"])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_ilq_iterates(oracle, name):
    """ILQSolver::Solve (src/ilq_solver.cpp:76-172): every logged operating point."""
    g = load(name)
    build, params = CASES[name]
    desc, x0_example = build()
    x0 = g["x0"]
    assert np.array_equal(np.asarray(x0_example, np.float32), x0[0])
    iters = int(g["ilq_iters"])
    B = x0.shape[0]
    h = abi.Handle(oracle, desc, params(max_solver_iters=iters), B)
    h.upload_x0(x0)
    h.solve_begin()
    # log iterate 0 = the initial rollout (ilq_solver.cpp:104-111)
    assert np.array_equal(h.download(abi.XS), g["ilq_xs"][:, 0])
    assert np.array_equal(h.download(abi.US), g["ilq_us"][:, 0])
    logged = np.ones(B, np.int32)
    for it in range(1, iters + 1):
        h.iterate(1)
        status, done = h.download(abi.STATUS), h.download(abi.ITERS)
        xs, us = h.download(abi.XS), h.download(abi.US)
        # an iteration is logged iff its linesearch succeeded (:140-165)
        newly = (done == it) & (status != abi.STATUS_LINESEARCH_FAILED)
        logged += newly
        for b in np.nonzero(newly)[0]:
            assert np.array_equal(xs[b], g["ilq_xs"][b, it]), (name, b, it)
            assert np.array_equal(us[b], g["ilq_us"][b, it]), (name, b, it)
    status = h.download(abi.STATUS)
    assert np.array_equal(logged, g["ilq_iterates"])
    assert np.array_equal((status != abi.STATUS_LINESEARCH_FAILED).astype(np.int32), g["ilq_success"])
    # final strategies: of the last logged iterate
    ok = g["ilq_success"] == 1
    assert np.array_equal(h.download(abi.PS)[ok], g["ilq_Ps"][ok])
    assert np.array_equal(h.download(abi.ALPHAS)[ok], g["ilq_alphas"][ok])
    assert ulp_close(h.download(abi.TOTAL_COSTS)[ok], g["ilq_costs"][ok], 2)
    h.close()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_lin_quad(oracle, name):
    """ComputeLinearization at the initial rollout and ComputeCostQuadraticization at the first
    accepted iterate (src/ilq_solver.cpp:437-490): every A, B, Q, l, R, r entry."""
    g = load(name)
    build, params = CASES[name]
    desc, _ = build()
    nlq = g["lq_A"].shape[0]
    h = abi.Handle(oracle, desc, params(max_solver_iters=1), nlq)
    h.upload_x0(g["x0"][:nlq])
    h.solve_begin()
    h.linearize_quadraticize()
    assert np.array_equal(h.download(abi.LIN_A), g["lq_A"])
    assert np.array_equal(h.download(abi.LIN_B), g["lq_B"])
    if params().linesearch:
        h.iterate(1)
        ok = h.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED
        h.linearize_quadraticize()
    else:
        # without the linesearch nothing re-quadraticizes after the step: what
        # ILQSolver::Quadraticization() holds after one iteration is still the initial rollout's
        ok = np.ones(nlq, bool)
    for what, key in ((abi.QUAD_Q, "lq_Q"), (abi.QUAD_L, "lq_l"), (abi.QUAD_R, "lq_R"),
                      (abi.QUAD_RGRAD, "lq_r")):
        assert np.array_equal(h.download(what)[ok], g[key][ok]), key
    assert ok.any()
    h.close()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_open_loop_iterates(oracle, name):
    """ILQSolver::Solve with SolverParams::open_loop, i.e. on LQOpenLoopSolver::Solve
    (src/lq_open_loop_solver.cpp:73-195): every logged operating point, final alphas, P = 0."""
    g = load(name)
    build, params = CASES[name]
    desc, _ = build()
    nol = g["ol_xs"].shape[0]
    iters = int(g["ol_iters"])
    h = abi.Handle(oracle, desc, params(max_solver_iters=iters, open_loop=1), nol)
    h.upload_x0(g["x0"][:nol])
    h.solve_begin()
    logged = np.ones(nol, np.int32)
    for it in range(1, iters + 1):
        h.iterate(1)
        status, done = h.download(abi.STATUS), h.download(abi.ITERS)
        xs, us = h.download(abi.XS), h.download(abi.US)
        newly = (done == it) & (status != abi.STATUS_LINESEARCH_FAILED)
        logged += newly
        for b in np.nonzero(newly)[0]:
            assert np.array_equal(xs[b], g["ol_xs"][b, it]), (name, b, it)
            assert np.array_equal(us[b], g["ol_us"][b, it]), (name, b, it)
    assert np.array_equal(logged, g["ol_iterates"])
    ok = g["ol_success"] == 1
    assert np.array_equal((h.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED).astype(np.int32),
                          g["ol_success"])
    assert np.array_equal(h.download(abi.ALPHAS)[ok], g["ol_alphas"][ok])
    assert float(g["ol_Ps_absmax"]) == 0.0 and np.all(h.download(abi.PS) == 0)
    h.close()


def receding_horizon_cases(lib, g, desc, params, from_plan=False):
    """For every fixture case: solve and write the solution back (or, from_plan, upload the
    reference's plan as the warm start), then SetUpNextRecedingHorizon; yields
    (case index, handle, new t0)."""
    nrh = g["rh_x0"].shape[1]
    for c, (t, runtime) in enumerate(g["rh_cases"]):
        h = abi.Handle(lib, desc, params(max_solver_iters=int(g["rh_iters"])), nrh, 0)
        h.upload_x0(g["x0"][:nrh])
        if from_plan:
            h.upload_warmstart(g["rh_plan_xs"], g["rh_plan_us"], g["rh_plan_Ps"], g["rh_plan_alphas"])
        else:
            h.solve_begin()
            h.solve(chunk=1)
            h.overwrite_solution()
        new_t0 = h.setup_next_receding_horizon(g["rh_x_meas"][c], float(t), float(runtime))
        yield c, h, new_t0
        h.close()


UNCONSTRAINED = ["roundabout_merging", "three_player_overtaking", "two_player_collision",
                 "two_player_collision_avoidance_reachability", "dubins_origin", "two_player_reachability",
                 "modified_air_3d", "modified_three_player_intersection", "skeleton",
                 "three_player_intersection_reachability"]


@pytest.mark.parametrize("name", UNCONSTRAINED)
def test_oracle_reproduces_reference_receding_horizon(oracle, name):
    """Problem::SetUpNextRecedingHorizon (src/problem.cpp:64-186): new initial state, shifted and
    extended operating point and strategies, new t0 -- bit for bit, for every unconstrained example
    (concatenated systems stitch the ego part of the nearest plan state in, TwoPlayerUnicycle4D
    takes that state whole; each subsystem kind integrates the zero-control extension)."""
    g = load(name)
    build, params = CASES[name]
    desc, _ = build()
    shifted = 0
    for from_plan in (False, True):
      for c, h, new_t0 in receding_horizon_cases(oracle, g, desc, params, from_plan):
        assert new_t0 == g["rh_t0"][c]
        assert np.array_equal(h.download(abi.X0), g["rh_x0"][c])
        for what, key in ((abi.WARM_XS, "rh_xs"), (abi.WARM_US, "rh_us"), (abi.WARM_PS, "rh_Ps"),
                          (abi.WARM_ALPHAS, "rh_alphas")):
            assert np.array_equal(h.download(what), g[key][c]), (c, key)
        shifted += int(np.any(g["rh_Ps"][c][:, -5:] == 0))
    assert shifted >= 6   # the plan really moved: trailing strategies are the zero extension


def final_plan(g, b):
    last = int(g["ilq_iterates"][b]) - 1
    return dict(xs=g["ilq_xs"][b, last], us=g["ilq_us"][b, last], Ps=g["ilq_Ps"][b], alphas=g["ilq_alphas"][b])


def long_plan(plan):
    """tests/golden/make_ref_golden.py:long_plan -- five made-up executed steps in front of `plan`."""
    return {k: np.concatenate([plan[k][:5] - np.float32(1.0), plan[k]]) for k in ("xs", "us", "Ps", "alphas")}


def plan_handle(lib, name, games, plan_t0, long):
    """A handle whose warm start is the final ILQ iterate of ref_<name>.npz for `games` games
    (five steps longer than the horizon if `long`), its OperatingPoint::t0 = plan_t0."""
    g = load(name)
    build, params = CASES[name]
    desc, _ = build()
    plans = [final_plan(g, b) for b in range(games)]
    if long:
        plans = [long_plan(p) for p in plans]
    desc.num_time_steps = plans[0]["xs"].shape[0]
    desc.initial_time = plan_t0
    h = abi.Handle(lib, desc, params(), games)
    h.upload_warmstart(*[np.stack([p[k] for p in plans]) for k in ("xs", "us", "Ps", "alphas")])
    return h


def integrate_plan_cases(lib, name):
    """Every MultiPlayerIntegrableSystem::Integrate(t0, t, ...) case of ref_integrate_<name>.npz:
    yields (tag, case, what the library returns, what the reference returned)."""
    gi = np.load(os.path.join(GOLDEN, f"ref_integrate_{name}.npz"))
    for tag, long in (("ip", False), ("ip_long", True)):
        for c, (plan_t0, t0, t) in enumerate(gi[f"{tag}_cases"]):
            h = plan_handle(lib, name, gi[f"{tag}_x_in"].shape[1], float(plan_t0), long)
            before = h.download(abi.WARM_XS)
            got = h.integrate_plan(gi[f"{tag}_x_in"][c], float(t0), float(t))
            assert np.array_equal(h.download(abi.WARM_XS), before)   # the plan is read only
            h.close()
            yield tag, c, got, gi[f"{tag}_x_out"][c]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_integrate(oracle, name):
    """MultiPlayerIntegrableSystem::Integrate(Time t0, Time t, x0, operating_point, strategies)
    (src/multi_player_integrable_system.cpp:54-83) -- what the receding-horizon simulator runs
    between solves -- bit for bit, including its time-index truncation on grid times."""
    n_cases = 0
    for tag, c, got, want in integrate_plan_cases(oracle, name):
        assert np.array_equal(got, want), (tag, c, np.abs(got - want).max())
        n_cases += 1
    assert n_cases == 9


def long_plan_receding_cases(lib):
    """Problem::OverwriteSolution(spliced plan) + SetUpNextRecedingHorizon as
    src/receding_horizon_simulator.cpp:105-109 runs them: the stored plan is five steps longer than
    the horizon; the first kNumTimeSteps steps of what the library leaves are the new problem."""
    gi = np.load(os.path.join(GOLDEN, "ref_integrate_roundabout_merging.npz"))
    for c, (t, runtime) in enumerate(gi["rhl_cases"]):
        h = plan_handle(lib, "roundabout_merging", gi["rhl_x_meas"].shape[1], float(gi["rhl_plan_t0"]), True)
        new_t0 = h.setup_next_receding_horizon(gi["rhl_x_meas"][c], float(t), float(runtime))
        yield c, h, new_t0, gi
        h.close()


def test_oracle_reproduces_reference_receding_horizon_from_spliced_plan(oracle):
    T = 100
    extended = 0
    for c, h, new_t0, gi in long_plan_receding_cases(oracle):
        assert new_t0 == gi["rhl_t0"][c]
        assert np.array_equal(h.download(abi.X0), gi["rhl_x0"][c])
        for what, key in ((abi.WARM_XS, "rhl_xs"), (abi.WARM_US, "rhl_us"), (abi.WARM_PS, "rhl_Ps"),
                          (abi.WARM_ALPHAS, "rhl_alphas")):
            assert np.array_equal(h.download(what)[:, :T], gi[key][c]), (c, key)
        extended += int(np.all(gi["rhl_Ps"][c][:, T - 1] == 0))
    assert 1 <= extended < 3   # one case fits in the stored plan, one runs past its end


def test_integrate_plan_argument_errors(oracle):
    """CHECK_GE(t, t0), CHECK_GE(t0, operating_point.t0) (:57-58), CHECK_LT(current_timestep,
    xs.size()) (:154) become error codes."""
    h = plan_handle(oracle, "air_3d", 2, 1.0, False)
    x = np.zeros((2, 3), np.float32)
    for t0, t in ((1.5, 1.4), (0.9, 1.2), (1.2, 11.0), (1.2, 11.5)):
        with pytest.raises(abi.IlqgError):
            h.integrate_plan(x, t0, t)
    assert h.integrate_plan(x, 1.2, 10.95).shape == (2, 3)
    h.close()


def test_receding_horizon_argument_errors(oracle):
    """Where the reference CHECK-fails (src/problem.cpp:68-71, :87) the ABI returns an error."""
    desc, _ = problems.roundabout_merging()
    h = abi.Handle(oracle, desc, problems.roundabout_params(), 2)
    x = problems.roundabout_x0_batch(2, 1)
    for t, runtime in ((-0.5, 0.1), (0.25, -0.1), (9.95, 0.2), (1.0, 0.1)):
        with pytest.raises(abi.IlqgError):
            h.setup_next_receding_horizon(x, t, runtime)
    desc_c, _ = problems.three_player_intersection()   # constrained: the reference aborts later
    hc = abi.Handle(oracle, desc_c, problems.three_player_intersection_params(), 1)
    with pytest.raises(abi.IlqgError):
        hc.setup_next_receding_horizon(problems.three_player_intersection_x0_batch(1, 1), 0.25, 0.1)


@pytest.mark.parametrize("name", ["three_player_intersection", "air_3d",
                                  "three_player_collision_avoidance_reachability", "one_player_reachability"])
def test_oracle_reproduces_reference_augmented_lagrangian(oracle, name):
    """AugmentedLagrangianSolver::Solve (src/augmented_lagrangian_solver.cpp:72-210): final
    operating point and strategies, multipliers, mu, NumIterates, success."""
    g = load(name)
    build, params = CASES[name]
    desc, _ = build()
    nal = g["al_xs"].shape[0]
    h = abi.Handle(oracle, desc, params(max_solver_iters=int(g["al_inner"])), nal)
    h.upload_x0(g["x0"][:nal])
    res = al.solve_augmented_lagrangian(h, max_solver_iters=int(g["al_outer"]),
                                        reset_lambdas=False, reset_mu=False, reset_problem=False,
                                        chunk=1)
    assert np.array_equal(res.iterates, g["al_iterates"])
    assert np.array_equal(res.success, g["al_success"])
    assert np.array_equal(res.xs, g["al_xs"])
    assert np.array_equal(res.us, g["al_us"])
    assert np.array_equal(res.Ps, g["al_Ps"])
    assert np.array_equal(res.alphas, g["al_alphas"])
    assert np.array_equal(h.download(abi.MU), g["al_mu"])
    lam = h.download(abi.LAMBDAS)
    for b in range(nal):
        assert np.array_equal(sorted_rows(lam[b]), g["al_lambdas_sorted"][b]), b
    h.close()


@pytest.mark.skipif(not (os.path.isdir("/root/reference") and
                         os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref",
                                                     "libilqg_ref.so"))),
                    reason="live reference build only exists in the builder container")
@pytest.mark.parametrize("which,name", list(enumerate([
    "three_player_intersection", "roundabout_merging", "air_3d", "three_player_overtaking", "two_player_collision",
    "two_player_collision_avoidance_reachability", "three_player_collision_avoidance_reachability",
    "one_player_reachability", "dubins_origin", "two_player_reachability", "modified_air_3d",
    "modified_three_player_intersection", "skeleton", "three_player_intersection_reachability"])))
def test_oracle_against_live_reference_on_fresh_seeds(oracle, which, name):
    """Same check on initial states that are NOT in the fixtures, calling the compiled reference
    directly (builder container only): the example's own initial state plus N(0, 0.1) on every
    state dimension, five ILQSolver iterations, final iterate and success flag bit for bit.
    `which` is the example's id in oracle/ref_driver.cpp."""
    from tests.golden import ref_lib as R
    ref = R.RefLibrary()
    build, params = CASES[name]
    desc, x0_example = build()
    rng = np.random.default_rng(1000 + which)
    x0 = (np.tile(x0_example, (4, 1)) + rng.normal(0, 0.1, size=(4, len(x0_example)))).astype(np.float32)
    p = params(max_solver_iters=5)
    h = abi.Handle(oracle, desc, p, 4)
    h.upload_x0(x0)
    h.solve_begin()
    h.solve(chunk=1)
    xs, us, status = h.download(abi.XS), h.download(abi.US), h.download(abi.STATUS)
    for b in range(4):
        r = ref.solve(which, R.ILQ, x0[b], R.RefParams.from_abi(p))
        assert np.array_equal(r["xs"][-1], xs[b]) and np.array_equal(r["us"][-1], us[b])
        assert r["success"] == int(status[b] != abi.STATUS_LINESEARCH_FAILED)
    h.close()


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"),
                    reason="the reference tree only exists in the builder container")
def test_reference_gtest_suite_passes_on_the_shim_build():
    """`make -C oracle ref-test`: the reference's own test/*.cpp (vendored gtest), unmodified,
    linked against the reference's own src/*.cpp compiled on oracle/ref_shim.  All of its tests
    pass, which is what qualifies the stand-ins for Eigen/glog/gflags (and with them the
    fixtures above) as "the reference run here"."""
    import subprocess
    repo = os.path.dirname(HERE)
    out = subprocess.run(["make", "-C", os.path.join(repo, "oracle"), "ref-test"], check=True,
                         capture_output=True, text=True).stdout
    assert "[  PASSED  ] 51 tests." in out, out[-2000:]
    assert "FAILED" not in out


def test_reset_solution_rewinds_the_receding_horizon_clock(oracle):
    """ILQG_RESET_SOLUTION = Problem::Initialize: zero plan and t0 back at the initial time, so the
    same receding-horizon call is accepted again (t0 must not precede the plan's t0)."""
    desc, _ = problems.roundabout_merging()
    h = abi.Handle(oracle, desc, problems.roundabout_params(max_solver_iters=1), 2)
    x = problems.roundabout_x0_batch(2, 1)
    h.upload_x0(x)
    assert h.setup_next_receding_horizon(x, 0.25, 0.1) > 0.25
    with pytest.raises(abi.IlqgError):          # the plan now starts after t = 0.25
        h.setup_next_receding_horizon(x, 0.05, 0.0)
    h.reset(h.RESET_SOLUTION)
    assert np.all(h.download(abi.WARM_XS) == 0)
    assert h.setup_next_receding_horizon(x, 0.05, 0.0) > 0.05


def test_polyline2_signed_distance_constraint_equals_the_references_class(oracle):
    """ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE against the reference's own class
    (src/polyline2_signed_distance_constraint.cpp:58-145; no example adds one, so the class is what there
    is to pin against): g and the augmented-Lagrangian derivatives at 64 seeded points per case --
    interior closest points and vertices, both orientations, with and without a multiplier -- bit for bit."""
    from tests import oracle_probes as probe
    g = np.load(os.path.join(GOLDEN, "ref_polyline_constraint.npz"))
    cases = len([k for k in g.files if k.endswith("_pts")])
    assert cases == 4
    vertices = 0
    for c in range(cases):
        pts, (threshold, keep_left, mu, _), xy, want = (g[f"case{c}_{k}"] for k in ("pts", "args", "xy", "out"))
        b = problems.DescBuilder(10, 0.1)
        b.add_player(1)
        b.d.xdim = 2
        b.state_constraint(0, abi.CONSTRAINT_POLYLINE2_SIGNED_DISTANCE, dims=(0, 1), value=float(threshold),
                           flag=int(keep_left), polyline=b.add_polyline([tuple(p) for p in pts.tolist()]))
        h = abi.Handle(oracle, b.build(), abi.SolverParams.defaults(), 1)
        for q, w in zip(xy, want):
            lam = float(w[1])
            assert np.float32(probe.evaluate_record(oracle, h, 0, q)) == w[0]
            hess, grad = probe.quadraticize_record(oracle, h, 0, q, lam, float(mu))
            got = np.array([grad[0], grad[1], hess[0, 0], hess[0, 1], hess[1, 1]], np.float32)
            assert np.array_equal(got, w[2:]), (c, q, got, w[2:])
            assert hess[0, 1] == hess[1, 0]
            vertices += probe.polyline_closest_point(oracle, h, 0, q)[1]
        h.close()
    assert vertices > 20  # (the vertex branch of Quadraticize is exercised, not only the interior one)
