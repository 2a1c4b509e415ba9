"""Pin the CPU oracle against the reference's OWN tests.

The reference C++ cannot be compiled here (Eigen3/glog/gflags absent), so each
test below transcribes a reference test -- same systems, same constructor
arguments, same tolerances -- with its file:line, and runs it against the
restatement in oracle/ilqg_oracle.cpp.  Nothing in this file touches the GPU.
"""
import math

import numpy as np
import pytest

from ilqgames_b200 import _abi as abi
from ilqgames_b200 import problems
from tests import oracle_probes as probe

kSmallNumber = 1e-4  # include/ilqgames/utils/types.h:116


# --------------------------------------------------------------------------
# test/test_line_segment2.cpp
# --------------------------------------------------------------------------
def test_line_segment_dies_if_degenerate(oracle):
    # test/test_line_segment2.cpp:52-54 (CHECK_GT(length_, kSmallNumber) -> error code here)
    rc, *_ = probe.segment_closest_point(oracle, (0, 0), (0, 0), (1, 1))
    assert rc == abi.OK - 1  # ILQG_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("query,is_endpoint,closest,ssd", [
    # test/test_line_segment2.cpp:57-102, segment (0,-1)-(0,1)
    ((1.0, -2.0), True, (0.0, -1.0), 2.0),
    ((1.0, 0.0), False, (0.0, 0.0), 1.0),
    ((1.0, 2.0), True, (0.0, 1.0), 2.0),
    ((-1.0, -2.0), True, (0.0, -1.0), -2.0),
    ((-1.0, 0.0), False, (0.0, 0.0), -1.0),
    ((-1.0, 2.0), True, (0.0, 1.0), -2.0),
])
def test_line_segment_closest_point_golden(oracle, query, is_endpoint, closest, ssd):
    rc, c, e, s = probe.segment_closest_point(oracle, (0.0, -1.0), (0.0, 1.0), query)
    assert rc == abi.OK
    assert e == is_endpoint
    assert np.allclose(c, closest, atol=1e-6)
    assert abs(s - ssd) < kSmallNumber


# --------------------------------------------------------------------------
# test/test_polyline2.cpp
# --------------------------------------------------------------------------
def _polyline_handle(lib, points):
    b = problems.DescBuilder(10, 0.1)
    b.add_player(1)
    b.d.xdim = 2
    b.add_polyline(points)
    return abi.Handle(lib, b.build(), abi.SolverParams.defaults(), 1)


@pytest.mark.parametrize("query,is_vertex,closest,ssd", [
    # test/test_polyline2.cpp:52-125, polyline (0,-1),(0,1),(2,1)
    ((1.0, -2.0), True, (0.0, -1.0), 2.0),
    ((0.5, 0.0), False, (0.0, 0.0), 0.25),
    ((1.5, 0.0), False, (1.5, 1.0), 1.0),
    ((3.0, 0.0), True, (2.0, 1.0), 2.0),
    ((-1.0, -2.0), True, (0.0, -1.0), -2.0),
    ((-1.0, 0.0), False, (0.0, 0.0), -1.0),
    ((-1.0, 2.0), True, (0.0, 1.0), -2.0),
    ((0.5, 2.0), False, (0.5, 1.0), -1.0),
    ((3.0, 2.0), True, (2.0, 1.0), -2.0),
])
def test_polyline_closest_point_golden(oracle, query, is_vertex, closest, ssd):
    h = _polyline_handle(oracle, [(0.0, -1.0), (0.0, 1.0), (2.0, 1.0)])
    c, v, _seg, s, _e = probe.polyline_closest_point(oracle, h, 0, query)
    assert v == is_vertex
    assert np.allclose(c, closest, atol=1e-6)
    assert abs(s - ssd) < kSmallNumber


# --------------------------------------------------------------------------
# test/test_lq_solver.cpp
# --------------------------------------------------------------------------
def lq_test_system(nominal=0.0, T=100, dt=0.1):
    """TwoPlayerPointMass1D + costs, test/test_lq_solver.cpp:143-177,227-248."""
    A = np.eye(2) + dt * np.array([[0.0, 1.0], [0.0, 0.0]])
    B1 = dt * np.array([[0.05], [1.0]])
    B2 = dt * np.array([[0.032], [0.11]])
    rel = 0.1  # kRelativeCostScaling
    Q = np.stack([np.eye(2), rel * np.eye(2)])
    l = np.stack([-nominal * np.ones(2), -rel * nominal * np.ones(2)])  # weight*(x - nominal), x=0
    # pairs sorted (0,0) (0,1) (1,0) (1,1)
    R = np.array([1.0, rel, rel, 1.0])
    r = np.array([-1.0, -rel, -rel, -1.0]) * nominal
    tile = lambda a: np.broadcast_to(a, (1, T) + a.shape).astype(np.float32).copy()
    return dict(A=tile(A), Bs=tile(np.concatenate([B1, B2], axis=1)), Q=tile(Q), l=tile(l),
                R=tile(R), r=tile(r)), (A, B1, B2, Q, R)


def lyapunov_iterations(A, B1, B2, Q1, Q2, R11, R12, R21, R22, iters=100):
    """SolveLyapunovIterations, test/test_lq_solver.cpp:72-109 (fp64 numpy)."""
    Z1, Z2 = Q1.copy(), Q2.copy()
    P1 = np.linalg.solve(R11 + B1.T @ Z1 @ B1, B1.T @ Z1 @ A)
    P2 = np.linalg.solve(R22 + B2.T @ Z2 @ B2, B2.T @ Z2 @ A)
    for _ in range(iters):
        o1, o2 = P1, P2
        P1 = np.linalg.solve(R11 + B1.T @ Z1 @ B1, B1.T @ Z1 @ (A - B2 @ o2))
        P2 = np.linalg.solve(R22 + B2.T @ Z2 @ B2, B2.T @ Z2 @ (A - B1 @ o1))
        F = A - B1 @ P1 - B2 @ P2
        Z1 = F.T @ Z1 @ F + P1.T @ R11 @ P1 + P2.T @ R12 @ P2 + Q1
        Z2 = F.T @ Z2 @ F + P1.T @ R21 @ P1 + P2.T @ R22 @ P2 + Q2
    return P1, P2


def solve_lq(lib, nominal):
    desc = problems.lq_only(100, 2, [1, 1], cross_pairs=[(0, 1), (1, 0)])
    h = abi.Handle(lib, desc, abi.SolverParams.defaults(), 1)
    lo = h.layout
    assert [(lo.pair_player[p], lo.pair_arg[p]) for p in range(lo.num_pairs)] == [
        (0, 0), (0, 1), (1, 0), (1, 1)]
    arrs, mats = lq_test_system(nominal)
    h.upload_lq(**arrs)
    h.lq_backward()
    return h.download(abi.LQ_PS)[0], h.download(abi.LQ_ALPHAS)[0], mats


def test_lq_feedback_matches_lyapunov_iterations(oracle):
    # LQFeedbackSolverTest.MatchesLyapunovIterations, test/test_lq_solver.cpp:292-317
    Ps, alphas, (A, B1, B2, Q, R) = solve_lq(oracle, 0.0)
    one = lambda v: np.array([[v]])
    P1, P2 = lyapunov_iterations(A, B1, B2, Q[0], Q[1], one(R[0]), one(R[1]), one(R[2]), one(R[3]))
    assert np.abs(P1 - Ps[0, 0:1]).max() < kSmallNumber
    assert np.abs(P2 - Ps[0, 1:2]).max() < kSmallNumber
    # SURVEY Appendix B known answers (probe values; the reference test only pins 1e-4)
    assert np.allclose(Ps[0, 0], [0.915024, 1.632025], atol=1e-4)
    assert np.allclose(Ps[0, 1], [0.0145362, 0.0206234], atol=1e-4)
    assert np.allclose(alphas[0], 0.0, atol=1e-7)
    # Strategy at k = T-1 is left at zero (strategy.h:64-70)
    assert np.all(Ps[-1] == 0) and np.all(alphas[-1] == 0)


def _rollout_cost(A, B1, B2, Q, l, R, r, Ps, alphas, x0, perturb=None):
    """Closed-loop cost of both players on the LQ game with u_i = -P_i x - alpha_i
    (ComputeStrategyCosts, src/compute_strategy_costs.cpp:61-105, linear dynamics)."""
    T = Ps.shape[0]
    x = x0.copy()
    costs = np.zeros(2)
    for k in range(T):
        u = -(Ps[k] @ x) - alphas[k]
        if perturb is not None:
            u = u + perturb[k]
        for i in range(2):
            costs[i] += 0.5 * x @ Q[i] @ x + l[i] @ x
            for j in range(2):
                Rij, rij = R[2 * i + j], r[2 * i + j]
                costs[i] += 0.5 * Rij * u[j] ** 2 + rij * u[j]
        x = A @ x + B1[:, 0] * u[0] + B2[:, 0] * u[1]
    return costs


@pytest.mark.parametrize("nominal", [0.0, 0.5])
def test_lq_feedback_is_local_nash(oracle, nominal):
    # LQFeedbackSolverTest.NashEquilibrium / ...WithLinearCostTerms,
    # test/test_lq_solver.cpp:319-345: no unilateral perturbation of a player's
    # alphas (|delta| <= 0.1) may lower that player's own closed-loop cost.
    Ps, alphas, (A, B1, B2, Q, R) = solve_lq(oracle, nominal)
    Ps, alphas = Ps.astype(np.float64), alphas.astype(np.float64)
    rel = 0.1
    l = np.stack([-nominal * np.ones(2), -rel * nominal * np.ones(2)])
    r = np.array([-1.0, -rel, -rel, -1.0]) * nominal
    x0 = np.ones(2)
    base = _rollout_cost(A, B1, B2, Q, l, R, r, Ps, alphas, x0)
    rng = np.random.default_rng(0)
    for _ in range(50):
        for i in range(2):
            k = rng.integers(0, 99)
            pert = np.zeros((100, 2))
            pert[k, i] = rng.uniform(-0.1, 0.1)
            c = _rollout_cost(A, B1, B2, Q, l, R, r, Ps, alphas, x0, pert)
            assert c[i] >= base[i] - 1e-6, (i, k, c[i], base[i])
    if nominal == 0.5:
        # alphas[0] for the nominal-0.5 case
        assert np.allclose(alphas[0], [-0.383023, -0.499542], atol=2e-4)


def solve_lq_open_loop(lib, nominal, x0):
    desc = problems.lq_only(100, 2, [1, 1], cross_pairs=[(0, 1), (1, 0)])
    h = abi.Handle(lib, desc, abi.SolverParams.defaults(open_loop=1), 1)
    arrs, mats = lq_test_system(nominal)
    h.upload_lq(**arrs)
    h.upload(abi.LQ_X0, np.asarray(x0, np.float32)[None])
    h.lq_backward()
    return h.download(abi.LQ_PS)[0], h.download(abi.LQ_ALPHAS)[0], h.download(abi.DELTA_XS)[0], mats


def _open_loop_cost(A, B1, B2, Q, l, R, r, alphas, x0, perturb=None):
    """ComputeStrategyCosts(..., open_loop = true), src/compute_strategy_costs.cpp:61-105: controls
    u_i = -alpha_i (zero operating point), T - 1 steps, state costs taken at the NEXT state
    (PlayerCost::EvaluateOffset, src/player_cost.cpp:175-192)."""
    T = alphas.shape[0]
    x = x0.copy()
    costs = np.zeros(2)
    for k in range(T - 1):
        u = -alphas[k]
        if perturb is not None:
            u = u - perturb[k]
        x = A @ x + B1[:, 0] * u[0] + B2[:, 0] * u[1]
        for i in range(2):
            costs[i] += 0.5 * x @ Q[i] @ x + l[i] @ x
            for j in range(2):
                costs[i] += 0.5 * R[2 * i + j] * u[j] ** 2 + r[2 * i + j] * u[j]
    return costs


def test_lq_open_loop_is_open_loop_nash(oracle):
    # LQOpenLoopSolverTest.NashEquilibrium, test/test_lq_solver.cpp:347-379 with
    # NumericalCheckLocalNashEquilibrium(..., open_loop = true)
    # (src/check_local_nash_equilibrium.cpp:60-135): alpha_i[k] -+ 0.1 for every player and step.
    nominal = 0.5
    x0 = np.ones(2)
    Ps, alphas, dxs, (A, B1, B2, Q, R) = solve_lq_open_loop(oracle, nominal, x0)
    assert np.all(Ps == 0)                       # :101-107 P is never touched
    assert np.all(alphas[-1] == 0)
    alphas = alphas.astype(np.float64)
    rel = 0.1
    l = np.stack([-nominal * np.ones(2), -rel * nominal * np.ones(2)])
    r = np.array([-1.0, -rel, -rel, -1.0]) * nominal
    base = _open_loop_cost(A, B1, B2, Q, l, R, r, alphas, x0)
    for i in range(2):
        for k in range(99):
            for sign in (-1.0, 1.0):
                pert = np.zeros((100, 2))
                pert[k, i] = sign * 0.1
                c = _open_loop_cost(A, B1, B2, Q, l, R, r, alphas, x0, pert)
                assert c[i] >= base[i] - 1e-6, (i, k, sign, c[i], base[i])
    # delta_xs is the optimal state trajectory: x*_0 = x0, x*_{k+1} = A x*_k - sum_i B_i alpha_i[k]
    x = x0.copy()
    for k in range(99):
        assert np.allclose(dxs[k], x, atol=2e-4)
        x = A @ x - B1[:, 0] * alphas[k, 0] - B2[:, 0] * alphas[k, 1]
    assert np.allclose(dxs[99], x, atol=2e-4)


def test_single_player_open_loop_equals_feedback_at_first_step(oracle):
    # SinglePlayerOpenLoopFeedback.TestSameSolution, test/test_lq_solver.cpp:381-436
    T, dt = 100, 0.1
    A = np.eye(2)
    A[0, 1] = dt
    B = dt * 0.41 * np.eye(2)
    tile = lambda a: np.broadcast_to(a, (1, T) + a.shape).astype(np.float32).copy()
    arrs = dict(A=tile(A), Bs=tile(B), Q=tile(np.eye(2)[None]), l=tile(np.zeros((1, 2))),
                R=tile(np.eye(2).reshape(-1)), r=tile(np.zeros(2)))
    x0 = np.ones(2, np.float32)
    out = {}
    for open_loop in (0, 1):
        h = abi.Handle(oracle, problems.lq_only(T, 2, [2]), abi.SolverParams.defaults(open_loop=open_loop), 1)
        h.upload_lq(**arrs)
        h.upload(abi.LQ_X0, x0[None])
        h.lq_backward()
        out[open_loop] = (h.download(abi.LQ_PS)[0], h.download(abi.LQ_ALPHAS)[0])
    # Strategy::operator()(k, dx, u_ref) = u_ref - P dx - alpha (strategy.h:73-76)
    u_ol = -out[1][1][0]                                  # (0, zero, zero)
    u_fb = -(out[0][0][0] @ x0) - out[0][1][0]            # (0, x0, zero)
    assert np.abs(u_ol - u_fb).max() < 0.01 * np.abs(u_fb).max()


# --------------------------------------------------------------------------
# test/test_quadraticization.cpp
# --------------------------------------------------------------------------
kInputDimension = 10
kPolyline = [(-2.0, -2.0), (0.5, 1.0), (2.0, 2.0)]
kBigPolyline = [(-200.0, -200.0), (0.5, 1.0), (200.0, 200.0)]

COST_CASES = {
    # name: (kind, kwargs, polyline points or None, is_constraint) -- constructor arguments
    # of test/test_quadraticization.cpp:205-332 for the record kinds on the hot path
    "QuadraticCost": (abi.COST_QUADRATIC, dict(dims=(-1,), weight=1.0, value=1.0), None, False),
    "QuadraticCostOneDim": (abi.COST_QUADRATIC, dict(dims=(3,), weight=2.0, value=0.7), None, False),
    "SemiquadraticCost": (abi.COST_SEMIQUADRATIC, dict(dims=(0,), weight=1.0, value=0.0, flag=1),
                          None, False),
    "QuadraticPolyline2Cost": (abi.COST_QUADRATIC_POLYLINE2, dict(dims=(0, 1), weight=1.0),
                               kPolyline, False),
    "SemiquadraticPolyline2Cost": (abi.COST_SEMIQUADRATIC_POLYLINE2,
                                   dict(dims=(0, 1), weight=1.0, value=0.5, flag=1), kBigPolyline,
                                   False),
    "ProximityCost": (abi.COST_PROXIMITY, dict(dims=(0, 1, 2, 3), weight=1.0, value=0.0), None,
                      False),
    "ProximityCostActive": (abi.COST_PROXIMITY, dict(dims=(0, 1, 2, 3), weight=1.0, value=20.0),
                            None, False),
    "Polyline2SignedDistanceCost": (abi.COST_POLYLINE2_SIGNED_DISTANCE,
                                    dict(dims=(0, 1), value=0.0, flag=1), kPolyline, False),
    "ProximityConstraint": (abi.CONSTRAINT_PROXIMITY, dict(dims=(0, 1, 2, 3), value=0.7, flag=0),
                            None, True),
    "SingleDimensionConstraint": (abi.CONSTRAINT_SINGLE_DIMENSION, dict(dims=(0,), value=1.0, flag=1),
                                  None, True),
    # test/test_quadraticization.cpp:323-327 (threshold 10, keep_left) and its mirror image
    "Polyline2SignedDistanceConstraint": (abi.CONSTRAINT_POLYLINE2_SIGNED_DISTANCE,
                                          dict(dims=(0, 1), value=10.0, flag=1), kPolyline, True),
    "Polyline2SignedDistanceConstraintRight": (abi.CONSTRAINT_POLYLINE2_SIGNED_DISTANCE,
                                               dict(dims=(0, 1), value=-0.5, flag=0), kPolyline, True),
    # kinds added with the widening of SURVEY 8 f4 (test/test_quadraticization.cpp has the same checks:
    # QuadraticDifferenceCostTest :210-214 with dims {0, 1} / {1, 2}, SignedDistanceCostTest :291-295)
    "SignedDistanceCost": (abi.COST_SIGNED_DISTANCE, dict(dims=(0, 1, 2, 3), value=5.0, flag=1), None, False),
    "SignedDistanceCostFlipped": (abi.COST_SIGNED_DISTANCE, dict(dims=(0, 1, 2, 3), value=5.0, flag=0), None,
                                  False),
    "QuadraticDifferenceCost": (abi.COST_QUADRATIC_DIFFERENCE, dict(dims=(0, 1, 1, 2), weight=1.0, flag=2), None,
                                False),
}


def _cost_handle(lib, kind, kw, poly, is_constraint):
    b = problems.DescBuilder(10, 0.1)
    b.add_player(1)
    b.d.xdim = kInputDimension
    kw = dict(kw)
    if poly is not None:
        kw["polyline"] = b.add_polyline(poly)
    (b.state_constraint if is_constraint else b.state_cost)(0, kind, **kw)
    return abi.Handle(lib, b.build(), abi.SolverParams.defaults(), 1)


@pytest.mark.parametrize("name", sorted(COST_CASES))
def test_quadraticization_matches_finite_differences(oracle64, name):
    # CheckQuadraticization, test/test_quadraticization.cpp:138-201: central differences with
    # step 1e-3, 20 random points with |entry| in [0.5, 5] and random signs; tolerance
    # max(0.15, 10% of the largest analytic entry).  Run on the fp64 build of the same
    # restatement so differences are formula errors, not rounding.  Constraints are checked
    # against EvaluateAugmentedLagrangian (:164-171) with lambda = 0 (kDefaultLambda) and mu = 10.
    kind, kw, poly, is_con = COST_CASES[name]
    h = _cost_handle(oracle64, kind, kw, poly, is_con)
    rng = np.random.default_rng(0)
    step = 1e-3
    lam, mu = (0.0, 10.0) if is_con else (0.0, 0.0)
    for _ in range(20):
        x = (rng.uniform(0.5, 5.0, kInputDimension) * rng.choice([-1.0, 1.0], kInputDimension))
        x = x.astype(np.float32)
        hess, grad = probe.quadraticize_record(oracle64, h, 0, x, lam, mu)
        g_num = np.zeros(kInputDimension)
        h_num = np.zeros((kInputDimension, kInputDimension))
        for i in range(kInputDimension):
            q = x.astype(np.float64).copy()
            q[i] = x[i] + step
            hi = probe.evaluate_record(oracle64, h, 0, q, lam, mu, with_al=is_con)
            _, ghi = probe.quadraticize_record(oracle64, h, 0, q, lam, mu)
            q[i] = x[i] - step
            lo = probe.evaluate_record(oracle64, h, 0, q, lam, mu, with_al=is_con)
            _, glo = probe.quadraticize_record(oracle64, h, 0, q, lam, mu)
            g_num[i] = 0.5 * (hi - lo) / step
            h_num[:, i] = 0.5 * (ghi - glo) / step
        tol_h = max(0.15, 0.1 * np.abs(hess).max())
        tol_g = max(0.15, 0.1 * np.abs(grad).max())
        assert np.abs(hess - h_num).max() < tol_h, (name, x)
        assert np.abs(grad - g_num).max() < tol_g, (name, x)


# --------------------------------------------------------------------------
# test/test_linearization.cpp
# --------------------------------------------------------------------------
def _dyn_handle(lib, which):
    if which == "three_player":
        desc, _ = problems.three_player_intersection()
    elif which == "roundabout":
        desc, _ = problems.roundabout_merging()
    elif which == "air3d":
        desc, _ = problems.air_3d()
    else:   # the dynamics added with the widening: Car5D, Dubins car, point mass, two-player unicycle
        desc, _ = {"car5d": problems.two_player_collision_avoidance_reachability, "dubins": problems.dubins_origin,
                   "point_mass": problems.modified_air_3d, "two_player_unicycle": problems.two_player_reachability}[which]()
    return abi.Handle(lib, desc, abi.SolverParams.defaults(), 1)


@pytest.mark.parametrize("which", ["three_player", "roundabout", "air3d", "car5d", "dubins", "point_mass",
                                   "two_player_unicycle"])
def test_linearization_matches_forward_differences(oracle, which):
    # CheckLinearization, test/test_linearization.cpp:142-196: A, B_i against forward
    # differences of Evaluate with h = 1e-3 (continuous-time Jacobian * dt + I), tol 1e-2,
    # 10 random points.  Covers Car6D, Unicycle4D, the concatenation and Air3D.
    h = _dyn_handle(oracle, which)
    n, M, dt, step = h.n, h.M, 0.1, 1e-3
    rng = np.random.default_rng(1)
    for _ in range(10):
        x = rng.uniform(-1.0, 1.0, n).astype(np.float32)
        u = rng.uniform(-1.0, 1.0, M).astype(np.float32)
        xs = np.tile(x, (1, h.T, 1))
        us = np.tile(u, (1, h.T, 1))
        h.upload_warmstart(xs=xs, us=us)
        h.linearize_quadraticize()
        A = h.download(abi.LIN_A)[0, 3]
        B = h.download(abi.LIN_B)[0, 3]
        f0, _ = probe.dynamics(oracle, h, x, u)
        A_num = np.eye(n)
        for j in range(n):
            q = x.copy()
            q[j] += step
            f1, _ = probe.dynamics(oracle, h, q, u)
            A_num[:, j] += dt * (f1 - f0) / step
        B_num = np.zeros((n, M))
        for j in range(M):
            q = u.copy()
            q[j] += step
            f1, _ = probe.dynamics(oracle, h, x, q)
            B_num[:, j] = dt * (f1 - f0) / step
        assert np.abs(A - A_num).max() < 1e-2
        assert np.abs(B - B_num).max() < 1e-2


# --------------------------------------------------------------------------
# test/test_player_cost.cpp
# --------------------------------------------------------------------------
def test_player_cost_evaluate_and_quadraticize(oracle):
    # PlayerCostTest, test/test_player_cost.cpp:60-122: QuadraticCost(weight, -1) on the state
    # and on both players' controls.  Evaluate = 0.5 w (|x|^2 + sum |u_i|^2); Quadraticize
    # gives w * I Hessians and w * x gradient.
    w, dim = 1.0, 4  # (reference uses 5; M = 2*dim must fit ILQG_MAX_UDIM = 8)
    b = problems.DescBuilder(4, 0.1)
    b.add_player(dim)
    b.add_player(dim)
    b.d.xdim = dim
    b.state_cost(0, abi.COST_QUADRATIC, dims=(-1,), weight=w, value=0.0)
    b.control_cost(0, 0, abi.COST_QUADRATIC, dims=(-1,), weight=w, value=0.0)
    b.control_cost(0, 1, abi.COST_QUADRATIC, dims=(-1,), weight=w, value=0.0)
    # player 1 needs its own control block to exist
    b.control_cost(1, 1, abi.COST_QUADRATIC, dims=(-1,), weight=w, value=0.0)
    params = abi.SolverParams.defaults()
    h = abi.Handle(oracle, b.build(), params, 1)
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, dim).astype(np.float32)
    us = rng.uniform(-1, 1, 2 * dim).astype(np.float32)
    h.upload_warmstart(xs=np.tile(x, (1, 4, 1)), us=np.tile(us, (1, 4, 1)))
    h.linearize_quadraticize()
    Q = h.download(abi.QUAD_Q)[0, 0, 0]
    l = h.download(abi.QUAD_L)[0, 0, 0]
    R = h.download(abi.QUAD_R)[0, 0]
    assert np.allclose(np.diag(Q), w, atol=kSmallNumber)
    assert abs(np.linalg.norm(Q) - w * math.sqrt(dim)) < kSmallNumber
    assert np.allclose(l, w * x, atol=kSmallNumber)
    lo = h.layout
    for p in range(lo.num_pairs):
        if lo.pair_player[p] != 0:
            continue
        Rp = R[lo.pair_R_offset[p]:lo.pair_R_offset[p] + dim * dim].reshape(dim, dim)
        assert np.allclose(np.diag(Rp), w, atol=kSmallNumber)
        assert abs(np.linalg.norm(Rp) - w * math.sqrt(dim)) < kSmallNumber
    # Evaluate: total cost of player 0 over T = 4 identical steps (TotalCosts sums over time)
    h.upload_x0(x[None])
    expected = 0.5 * w * (x @ x + us @ us)
    assert abs(probe.evaluate_record(oracle, h, 0, x) - 0.5 * w * (x @ x)) < kSmallNumber
    total = (probe.evaluate_record(oracle, h, 0, x) + probe.evaluate_record(oracle, h, 1, us[:dim])
             + probe.evaluate_record(oracle, h, 2, us[dim:]))
    assert abs(total - expected) < kSmallNumber


# --------------------------------------------------------------------------
# index quirks that must be reproduced bit-exactly
# --------------------------------------------------------------------------
def test_lambda_time_index_quirk(oracle):
    desc, _ = problems.three_player_intersection(num_time_steps=200)
    h = abi.Handle(oracle, desc, abi.SolverParams.defaults(), 1)
    idx = np.array(h.layout.lambda_index[:200])
    off = [k for k in range(200) if idx[k] != k]
    assert off == [43, 81, 86, 91, 162, 167, 172, 177, 182, 187]
    assert all(idx[k] == k - 1 for k in off)
    # independent restatement of Constraint::TimeIndex(RelativeTime(kk)) in Python doubles
    assert [int((float(k) * 0.1 - 0.0) / 0.1) for k in range(200)] == idx.tolist()
