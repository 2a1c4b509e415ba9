"""Generate tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN SOURCES.

`make -C oracle ref` compiles /root/reference/src/*.cpp (unmodified, in place) against the
Eigen/glog/gflags stand-ins in oracle/ref_shim into oracle/_ref/libilqg_ref.so; this script calls
the reference's ILQSolver::Solve / AugmentedLagrangianSolver::Solve on the reference's own example
Problems through oracle/ref_driver.cpp and stores what they return.  /root/reference only exists
in the builder container, so the outputs are committed as fixtures; tests/test_ref_pins.py checks
the CPU oracle against them (bit for bit) and tests/test_gpu_parity.py the CUDA path (tolerance).

What the fixtures pin: control flow, indexing, term order, linesearch, convergence test,
multiplier updates, every scalar formula of the reference.  What they do not: the rounding of
Eigen's own kernels (dense products, Householder QR), which the stand-in computes in plain
dot-product order.

Re-run (builder container only):  python tests/golden/make_ref_golden.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from ilqgames_b200 import problems  # noqa: E402
from tests.golden import ref_lib as R  # noqa: E402

ILQ_ITERS = 6       # ILQSolver::Solve cap for the per-iterate fixtures
OL_ITERS = 4        # same with SolverParams::open_loop
RH_ITERS = 2        # ILQ iterations before SetUpNextRecedingHorizon
# (t0, planner_runtime) pairs; times on the 0.1 s grid CHECK-fail in the reference (problem.cpp:87)
SPLICE_T0S = [0.3, 1.2, 0.0]   # start times of the spliced-in horizon (tests/cpp/host_api_test.cpp uses the same)
RH_CASES = [(0.25, 0.1), (0.33, 0.25), (1.02, 0.1), (0.55, 0.0), (2.07, 0.5)]
AL_INNER, AL_OUTER = 10, 40   # unconstrained_solver_max_iters, AL NumIterates cap
# MultiPlayerIntegrableSystem::Integrate(t0, t, ...): (plan t0, t0, t).  t0 at the plan's start (no
# IntegrateToNextTimeStep), both times inside one step, t == t0, times on the 0.1 s grid (where
# relative_t / kTimeStep truncates one step low), a plan that does not start at zero.
IP_CASES = [(0.0, 0.0, 0.25), (0.0, 0.13, 0.52), (0.0, 0.30, 0.34), (0.0, 0.4, 0.4), (0.0, 1.07, 2.5),
            (0.0, 0.2, 0.7), (1.3, 1.51, 2.25)]
# the same on a plan five steps longer than the horizon, as SolutionSplicer hands over
IP_LONG_CASES = [(0.5, 0.77, 1.9), (0.5, 10.2, 10.93)]
# SetUpNextRecedingHorizon from the long plan (plan t0 = 0.5): (t, planner_runtime); the first fits in
# the stored plan, the second runs past its end and is extended with zero controls
RHL_T0 = 0.5
RHL_CASES = [(0.62, 0.1), (0.97, 0.45), (1.33, 0.2)]
IP_GAMES = 3

CASES = {
    # name: (reference problem id, descriptor builder, params builder, x0 batch)
    "three_player_intersection": (R.INTERSECTION, problems.three_player_intersection,
                                  problems.three_player_intersection_params,
                                  lambda: problems.three_player_intersection_x0_batch(12, 1024)),
    "roundabout_merging": (R.ROUNDABOUT, problems.roundabout_merging, problems.roundabout_params,
                           lambda: problems.roundabout_x0_batch(8, 4096)),
    "air_3d": (R.AIR3D, problems.air_3d, problems.air_3d_params,
               lambda: problems.air_3d_x0_grid(4)[:12]),
    # a fourth example of the reference that needs no record kind beyond the three above (n = 18)
    "three_player_overtaking": (R.OVERTAKING, problems.three_player_overtaking,
                                problems.three_player_overtaking_params,
                                lambda: problems.three_player_overtaking_x0_batch(8, 18)),
    # FinalTimeCost-wrapped goal costs (records with a time gate), n = 12, two players
    "two_player_collision": (R.COLLISION, problems.two_player_collision, problems.two_player_collision_params,
                             lambda: problems.two_player_collision_x0_batch(8, 12)),
    # SinglePlayerCar5D dynamics and SignedDistanceCost, both players max-over-time, n = 10
    "two_player_collision_avoidance_reachability": (
        R.REACHABILITY2, problems.two_player_collision_avoidance_reachability,
        problems.two_player_collision_avoidance_reachability_params,
        lambda: problems.two_player_collision_avoidance_reachability_x0_batch(8, 10)),
    # ExtremeValueCost (a group of records), box constraints on the controls, n = 15
    "three_player_collision_avoidance_reachability": (
        R.REACHABILITY3, problems.three_player_collision_avoidance_reachability,
        problems.three_player_collision_avoidance_reachability_params,
        lambda: problems.three_player_collision_avoidance_reachability_x0_batch(8, 15)),
    # a single player: SinglePlayerDubinsCar, n = 3, m = 1
    "one_player_reachability": (R.REACHABILITY1, problems.one_player_reachability,
                                problems.one_player_reachability_params,
                                lambda: problems.one_player_reachability_x0_batch(8, 3)),
    # QuadraticDifferenceCost; the executable's parameters switch the linesearch off
    "dubins_origin": (R.DUBINS_ORIGIN, problems.dubins_origin, problems.dubins_origin_params,
                      lambda: problems.dubins_origin_x0_batch(8, 6)),
    # TwoPlayerUnicycle4D: a coupled (non-concatenated) system besides Air3D, max / min over time
    "two_player_reachability": (R.REACHABILITY_2P, problems.two_player_reachability,
                                problems.two_player_reachability_params,
                                lambda: problems.two_player_reachability_x0_batch(8, 4)),
    # SinglePlayerPointMass2D (linear dynamics), cost weights of -1e6 / +1e6
    "modified_air_3d": (R.MODIFIED_AIR3D, problems.modified_air_3d, problems.modified_air_3d_params,
                        lambda: problems.modified_air_3d_x0_batch(8, 8)),
}


_SOURCE_DUMP = {}


def source_descriptor(tag: str):
    """(desc, x0) of a reference example as include/ilqgames/**'s DescribeProblem emits it from the
    example's OWN source, compiled unchanged (tests/cpp/host_api_test.cpp, drop-in build): for
    examples made of record kinds that already exist nobody has to transcribe constants by hand."""
    if not _SOURCE_DUMP:
        import pathlib
        import tempfile
        from ilqgames_b200 import _abi as abi
        from tests import test_host_api as T
        tmp = pathlib.Path(tempfile.mkdtemp())
        exe = T._build(tmp, os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so"), dropin=True)
        _SOURCE_DUMP.update(T._run(exe, tmp))
        _SOURCE_DUMP["__abi"] = abi
    abi = _SOURCE_DUMP["__abi"]
    desc = abi.ProblemDesc.from_buffer_copy(_SOURCE_DUMP["desc_" + tag].tobytes())
    return desc, _SOURCE_DUMP["x0_" + tag].copy()


def generic_params(initial_alpha_scaling, convergence_tolerance, expected_decrease_fraction):
    def make(**overrides):
        from ilqgames_b200 import _abi as abi
        base = dict(max_backtracking_steps=100, linesearch=1, expected_decrease_fraction=expected_decrease_fraction,
                    initial_alpha_scaling=initial_alpha_scaling, convergence_tolerance=convergence_tolerance)
        base.update(overrides)
        return abi.SolverParams.defaults(**base)
    return make


def generic_x0_batch(tag: str, batch: int, seed: int):
    """The example's own initial state plus N(0, 0.05) on every state dimension."""
    _, x0 = source_descriptor(tag)
    rng = np.random.default_rng(seed)
    return (np.tile(x0, (batch, 1)) + rng.normal(0, 0.05, size=(batch, len(x0)))).astype(np.float32)


# examples whose descriptor comes from their own source: name -> (reference id, dump tag, SolverParams of
# the example's executable: initial_alpha_scaling, convergence_tolerance, expected_decrease)
FROM_SOURCE = {
    "modified_three_player_intersection": (R.MODIFIED_INTERSECTION, "modified_intersection", (1.0, 1.0, 0.9)),
    "skeleton": (R.SKELETON, "skeleton", (0.25, 0.01, 0.1)),
    "three_player_intersection_reachability": (R.INTERSECTION_REACHABILITY, "intersection_reachability", (0.1, 0.01, 0.1)),
}


def sorted_rows(a: np.ndarray) -> np.ndarray:
    """Multipliers as a set of per-constraint rows: the reference keeps control constraints in an
    unordered_multimap, so their order is an implementation detail."""
    if a.shape[0] == 0:
        return a
    return a[np.lexsort(a.T[::-1])]


for _name, (_which, _tag, _p) in FROM_SOURCE.items():
    CASES[_name] = (_which, (lambda t=_tag: source_descriptor(t)), generic_params(*_p),
                    (lambda t=_tag: generic_x0_batch(t, 8, 7)))


def run_case(ref: R.RefLibrary, name: str):
    which, build, params, x0f = CASES[name]
    _, x0_example = build()
    x0 = x0f().astype(np.float32)
    x0[0] = x0_example           # instance 0 = the example's own initial state
    assert np.array_equal(ref.x0(which), x0[0]), "descriptor x0 differs from the reference example"
    B = x0.shape[0]
    n, M, N, T, nc = ref.dims(which)
    p = params(max_solver_iters=ILQ_ITERS)
    rp = R.RefParams.from_abi(p)
    out = {"x0": x0, "ilq_iters": np.int32(ILQ_ITERS)}
    if name in FROM_SOURCE:
        # the test side has no hand-written builder for these: it reads the descriptor from here
        desc, _ = build()
        out["desc_bytes"] = np.frombuffer(bytes(desc), np.uint8).copy()
        out["solver_params"] = np.array(FROM_SOURCE[name][2], np.float64)
    xs = np.full((B, ILQ_ITERS + 1, T, n), np.nan, np.float32)
    us = np.full((B, ILQ_ITERS + 1, T, M), np.nan, np.float32)
    Ps = np.zeros((B, T, M, n), np.float32)
    alphas = np.zeros((B, T, M), np.float32)
    costs = np.zeros((B, N), np.float32)
    iterates = np.zeros(B, np.int32)
    success = np.zeros(B, np.int32)
    for b in range(B):
        r = ref.solve(which, R.ILQ, x0[b], rp, max_log=ILQ_ITERS + 1)
        k = r["iterates"]
        xs[b, :k], us[b, :k] = r["xs"], r["us"]
        Ps[b], alphas[b], costs[b] = r["Ps"], r["alphas"], r["costs"]
        iterates[b], success[b] = k, r["success"]
    out.update(ilq_xs=xs, ilq_us=us, ilq_Ps=Ps, ilq_alphas=alphas, ilq_costs=costs,
               ilq_iterates=iterates, ilq_success=success)

    # linearization at the initial rollout, quadraticization at the first accepted iterate
    nlq = 2
    lq = [ref.lin_quad(which, x0[b], rp) for b in range(nlq)]
    for key in ("A", "B", "Q", "l", "R", "r"):
        out[f"lq_{key}"] = np.stack([d[key] for d in lq])

    # SolverParams::open_loop: ILQSolver on LQOpenLoopSolver (ilq_solver.h:76-81)
    nol = min(6, B)
    pol = params(max_solver_iters=OL_ITERS, open_loop=1)
    rpol = R.RefParams.from_abi(pol)
    ol = [ref.solve(which, R.ILQ, x0[b], rpol, max_log=OL_ITERS + 1) for b in range(nol)]
    ol_xs = np.full((nol, OL_ITERS + 1, T, n), np.nan, np.float32)
    ol_us = np.full((nol, OL_ITERS + 1, T, M), np.nan, np.float32)
    for b, r in enumerate(ol):
        ol_xs[b, :r["iterates"]], ol_us[b, :r["iterates"]] = r["xs"], r["us"]
    out.update(ol_iters=np.int32(OL_ITERS), ol_xs=ol_xs, ol_us=ol_us,
               ol_alphas=np.stack([r["alphas"] for r in ol]),
               ol_Ps_absmax=np.float32(max(np.abs(r["Ps"]).max() for r in ol)),
               ol_iterates=np.array([r["iterates"] for r in ol], np.int32),
               ol_success=np.array([r["success"] for r in ol], np.int32))

    if nc == 0:
        # Problem::SetUpNextRecedingHorizon (src/problem.cpp:127-186) after an ILQSolver::Solve whose
        # final iterate was written back with OverwriteSolution; measured states = plan + noise
        nrh = 3
        prh = params(max_solver_iters=RH_ITERS)
        rprh = R.RefParams.from_abi(prh)
        rng = np.random.default_rng(3)
        sol = [ref.solve(which, R.ILQ, x0[b], rprh, max_log=RH_ITERS + 1) for b in range(nrh)]
        plans = [r["xs"][-1] for r in sol]
        # the plan SetUpNextRecedingHorizon starts from (lets a test skip the solve)
        out.update(rh_plan_xs=np.stack(plans), rh_plan_us=np.stack([r["us"][-1] for r in sol]),
                   rh_plan_Ps=np.stack([r["Ps"] for r in sol]),
                   rh_plan_alphas=np.stack([r["alphas"] for r in sol]))
        rh = {k: [] for k in ("x_meas", "x0", "xs", "us", "Ps", "alphas", "t0")}
        for (t, runtime) in RH_CASES:
            k = int(t / 0.1)
            xm = np.stack([plans[b][k] for b in range(nrh)]) + rng.normal(0, 0.05, size=(nrh, n)).astype(np.float32)
            res = [ref.receding_horizon(which, x0[b], rprh, xm[b], t, runtime) for b in range(nrh)]
            rh["x_meas"].append(xm)
            for key in ("x0", "xs", "us", "Ps", "alphas"):
                rh[key].append(np.stack([r[key] for r in res]))
            assert len({r["t0"] for r in res}) == 1
            rh["t0"].append(res[0]["t0"])
        out.update(rh_iters=np.int32(RH_ITERS), rh_cases=np.array(RH_CASES, np.float64),
                   **{f"rh_{k}": np.array(v) for k, v in rh.items()})

    if nc > 0:
        nal = 6
        pa = params(max_solver_iters=AL_INNER)
        rpa = R.RefParams.from_abi(pa)
        rpa.unconstrained_solver_max_iters = AL_INNER
        rpa.max_solver_iters = AL_OUTER
        res = [ref.solve(which, R.AL, x0[b], rpa, max_log=AL_OUTER + AL_INNER + 2) for b in range(nal)]
        out.update(al_inner=np.int32(AL_INNER), al_outer=np.int32(AL_OUTER),
                   al_xs=np.stack([r["xs"][-1] for r in res]),
                   al_us=np.stack([r["us"][-1] for r in res]),
                   al_Ps=np.stack([r["Ps"] for r in res]),
                   al_alphas=np.stack([r["alphas"] for r in res]),
                   al_lambdas_sorted=np.stack([sorted_rows(r["lambdas"]) for r in res]),
                   al_mu=np.array([r["mu"] for r in res], np.float32),
                   al_iterates=np.array([r["iterates"] for r in res], np.int32),
                   al_success=np.array([r["success"] for r in res], np.int32))
    return out


def long_plan(plan: dict, t0: float) -> dict:
    """A plan five time steps longer than the horizon: five made-up executed steps (the plan's
    first five, displaced so that no nearest-state search lands on them) in front of `plan`."""
    out = {k: np.concatenate([plan[k][:5] - np.float32(1.0), plan[k]]) for k in ("xs", "us", "Ps", "alphas")}
    out["t0"] = t0
    return out


def final_plan(g, b: int) -> dict:
    last = int(g["ilq_iterates"][b]) - 1
    return dict(xs=g["ilq_xs"][b, last], us=g["ilq_us"][b, last], Ps=g["ilq_Ps"][b], alphas=g["ilq_alphas"][b])


def run_integrate(ref: R.RefLibrary, name: str, g) -> dict:
    """Dynamics()->Integrate(t0, t, x0, plan) under the final ILQ iterate of ref_<name>.npz, from
    measured states = plan state at t0 + noise; and, for the unconstrained problem,
    OverwriteSolution(long plan) + SetUpNextRecedingHorizon."""
    which = CASES[name][0]
    n = ref.dims(which)[0]
    rng = np.random.default_rng(11)
    out = {"ip_cases": np.array(IP_CASES, np.float64), "ip_long_cases": np.array(IP_LONG_CASES, np.float64)}
    for tag, cases, make in (("ip", IP_CASES, lambda p, t0: dict(p, t0=t0)), ("ip_long", IP_LONG_CASES, long_plan)):
        xin, xout = [], []
        for (plan_t0, t0, t) in cases:
            rows_in, rows_out = [], []
            for b in range(IP_GAMES):
                plan = make(final_plan(g, b), plan_t0)
                k = int((t0 - plan_t0) / 0.1)
                x = (plan["xs"][k] + rng.normal(0, 0.05, size=n)).astype(np.float32)
                rows_in.append(x)
                rows_out.append(ref.integrate(which, plan, x, t0, t))
            xin.append(np.stack(rows_in))
            xout.append(np.stack(rows_out))
        out[f"{tag}_x_in"], out[f"{tag}_x_out"] = np.stack(xin), np.stack(xout)
    if ref.dims(which)[4] == 0:
        rh = {k: [] for k in ("x_meas", "x0", "xs", "us", "Ps", "alphas", "t0")}
        for (t, runtime) in RHL_CASES:
            res, xm = [], []
            for b in range(IP_GAMES):
                plan = long_plan(final_plan(g, b), RHL_T0)
                k = int((t - RHL_T0) / 0.1)
                xm.append((plan["xs"][k] + rng.normal(0, 0.05, size=n)).astype(np.float32))
                res.append(ref.receding_from_plan(which, plan, xm[-1], t, runtime))
            rh["x_meas"].append(np.stack(xm))
            for key in ("x0", "xs", "us", "Ps", "alphas"):
                rh[key].append(np.stack([r[key] for r in res]))
            assert len({r["t0"] for r in res}) == 1
            rh["t0"].append(res[0]["t0"])
        out.update(rhl_cases=np.array(RHL_CASES, np.float64), rhl_plan_t0=np.float64(RHL_T0),
                   **{f"rhl_{k}": np.array(v) for k, v in rh.items()})
    return out


def check_polylines(ref: R.RefLibrary):
    """The lane / target polylines of problems.py equal the reference's, bit for bit."""
    import math
    F = np.float32
    off, wedge = F(math.pi / 2 * 0.5), F(math.pi)
    out = {}
    for i, dist in enumerate((25.0, 10.0, 25.0, 10.0)):
        ang = F(float(off) + i * 2.0 * math.pi / 4.0)
        mine = np.asarray(problems.roundabout_lane_center(ang, F(ang + wedge), dist), np.float32)
        theirs = ref.roundabout_lane(ang, F(ang + wedge), dist)
        assert np.array_equal(mine, theirs), (i, np.abs(mine - theirs).max())
        out[f"roundabout_lane_{i}"] = theirs
    mine = np.asarray(problems.draw_circle((0.0, 0.0), 5.0, 10), np.float32)
    theirs = ref.draw_circle(0.0, 0.0, 5.0, 10)
    assert np.array_equal(mine, theirs), np.abs(mine - theirs).max()
    out["air_3d_circle"] = theirs
    return out


POLYLINE_CONSTRAINT_CASES = [
    # (polyline, threshold, keep_left, mu, lambda_step): the reference test's polyline and arguments
    # (test/test_quadraticization.cpp:323-327), its mirror image, a lane boundary of the intersection
    # example (src/three_player_intersection_example.cpp:229-236: lane2, half width 2.5) with a
    # non-zero multiplier
    ([(-2.0, -2.0), (0.5, 1.0), (2.0, 2.0)], 10.0, True, 10.0, 0.0),
    ([(-2.0, -2.0), (0.5, 1.0), (2.0, 2.0)], -0.5, False, 10.0, 0.0),
    ([(-10.0, 1000.0), (-10.0, 18.0), (-9.5, 15.0), (-9.0, 14.0), (-7.0, 12.5), (-4.0, 12.0), (1000.0, 12.0)],
     2.5, False, 10.0, 0.3),
    ([(-10.0, 1000.0), (-10.0, 18.0), (-9.5, 15.0), (-9.0, 14.0), (-7.0, 12.5), (-4.0, 12.0), (1000.0, 12.0)],
     -2.5, True, 25.0, 0.07),
]


def run_polyline_constraint(ref: R.RefLibrary):
    """The reference's own Polyline2SignedDistanceConstraint class (no example adds one) at seeded points
    around each polyline -- interior closest points, vertices, both sides -- for ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE."""
    out = {}
    rng = np.random.default_rng(11)
    for c, (pts, threshold, keep_left, mu, lam_step) in enumerate(POLYLINE_CONSTRAINT_CASES):
        pts = np.asarray(pts, np.float32)
        inner = pts[np.abs(pts).max(axis=1) < 500.0]
        centre, spread = inner.mean(axis=0), np.maximum(np.ptp(inner, axis=0), 4.0)
        xy = (centre + rng.uniform(-1.0, 1.0, size=(64, 2)) * spread).astype(np.float32)
        out[f"case{c}_pts"] = pts
        out[f"case{c}_args"] = np.asarray([threshold, float(keep_left), mu, lam_step], np.float32)
        out[f"case{c}_xy"] = xy
        out[f"case{c}_out"] = ref.polyline_constraint(pts, threshold, keep_left, mu, lam_step, xy)
    return out


if __name__ == "__main__":
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "ref"], check=True,
                   stdout=subprocess.DEVNULL)
    ref = R.RefLibrary()
    np.savez_compressed(os.path.join(HERE, "ref_polylines.npz"), **check_polylines(ref))
    np.savez_compressed(os.path.join(HERE, "ref_polyline_constraint.npz"), **run_polyline_constraint(ref))
    if "--polyline-constraint-only" in sys.argv:
        sys.exit(0)
    # SolutionSplicer::Splice on synthetic logs (new horizon inside / beyond the keep window / at t0)
    splice = {}
    for c, new_t0 in enumerate(SPLICE_T0S):
        for k, v in ref.splice(new_t0).items():
            splice[f"splice{c}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_splice.npz"), **splice)
    only_integrate = "--integrate-only" in sys.argv   # keep the committed solver fixtures, add ref_integrate_*
    for name in CASES:
        if "--only" in sys.argv and name != sys.argv[sys.argv.index("--only") + 1]:
            continue
        if only_integrate:
            out = dict(np.load(os.path.join(HERE, f"ref_{name}.npz")))
        else:
            out = run_case(ref, name)
            np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
        np.savez_compressed(os.path.join(HERE, f"ref_integrate_{name}.npz"), **run_integrate(ref, name, out))
        print("wrote", name, "iterates", out["ilq_iterates"].tolist(), "success",
              out["ilq_success"].tolist(),
              "AL iterates", out.get("al_iterates", np.zeros(0)).tolist())
