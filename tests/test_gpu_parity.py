"""GPU parity tests proper: the sm_100a library, called through the C ABI, against the CPU oracle
on identical seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full
batch size -- through size-independent properties.

Tolerances (fp32 path; SURVEY.md section 7 "hard parts"):
  * integer / index outputs (status, iteration and rollout counts, extreme-time indices,
    lambda index map): bit-exact;
  * one stage on identical inputs (records, LQ solution, rollout): 1e-4 of the array's max-abs
    scale.  Measured on B200: <= 2e-6, and the fp64 build of the oracle shows the CUDA path and
    the fp32 oracle are equally far (same order) from the fp64 answer;
  * multi-iteration trajectories: instances whose control flow (Armijo decisions) matches are
    compared at 2e-2 absolute; an Armijo flip on a knife edge is legitimate
    and is bounded by a minimum matching fraction instead.
"""
import os

import numpy as np
import pytest

from ilqgames_b200 import _abi as abi
from ilqgames_b200 import problems
from tests.test_oracle_pins import lq_test_system, lyapunov_iterations

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAGE_TOL = 1e-4
# Fraction of the well-posed games on which the CUDA path must take exactly the oracle's Armijo
# decisions (status, iteration count, backtrack count).  Measured on B200 in round 2 with the counts
# every test now reports (gpurun_out/parity_counts.jsonl -> profiles/r02_parity_counts.md): the
# thresholds sit a margin below the measured fractions.
FLOW_MIN = 0.85


def flow_ok(flow, stable):
    """At least FLOW_MIN of the stable games follow the oracle's control flow; with a handful of
    stable games one stray knife-edge decision is allowed (1 of 6 is already below 85 %)."""
    n = int(stable.sum())
    misses = int((stable & ~flow).sum())
    return misses <= max(1, int((1.0 - FLOW_MIN) * n))

CONFIGS = {
    "three_player_intersection": (problems.three_player_intersection,
                                  problems.three_player_intersection_params,
                                  lambda b: problems.three_player_intersection_x0_batch(b, 1024)),
    "roundabout_merging": (problems.roundabout_merging, problems.roundabout_params,
                           lambda b: problems.roundabout_x0_batch(b, 4096)),
    # 8 x 8 grid of relative positions (round 1 used a 4 x 4 grid: "batch 36" was really 16 games)
    "air_3d": (problems.air_3d, problems.air_3d_params, lambda b: problems.air_3d_x0_grid(8)[:b]),
    # widening: a fourth example of the reference, same record kinds, n = 18
    "three_player_overtaking": (problems.three_player_overtaking, problems.three_player_overtaking_params,
                                lambda b: problems.three_player_overtaking_x0_batch(b, 18)),
}
HEADLINE = ["air_3d", "roundabout_merging", "three_player_intersection"]


def _fixture_config(name):
    """The other examples of the reference (row f4): descriptor and parameters as tests/test_ref_pins.py
    builds them, initial states = the ones the reference fixture was generated on."""
    from tests.test_ref_pins import CASES
    build, params = CASES[name]
    x0 = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))["x0"]
    return (lambda num_time_steps=100: build()), params, (lambda b: x0[:b])


WIDENED = ["two_player_collision", "two_player_collision_avoidance_reachability",
           "three_player_collision_avoidance_reachability", "one_player_reachability", "dubins_origin",
           "two_player_reachability", "modified_air_3d", "modified_three_player_intersection", "skeleton",
           "three_player_intersection_reachability"]
for _n in WIDENED:
    CONFIGS[_n] = _fixture_config(_n)


COUNTS = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out", "parity_counts.jsonl")


def report(test, **numbers):
    """Every masked comparison says how much it compared: printed (pytest -s / -rP) and appended to
    gpurun_out/parity_counts.jsonl when that directory exists (the GPU box's scratch)."""
    import json
    line = json.dumps({"test": test, **{k: (float(v) if isinstance(v, (float, np.floating)) else int(v))
                                        for k, v in numbers.items()}})
    print("[parity]", line)
    if os.path.isdir(os.path.dirname(COUNTS)):
        with open(COUNTS, "a") as f:
            f.write(line + "\n")


def tame(ref, limit=1e6):
    """Per-instance mask: the oracle's values are finite and below `limit`.  Some sampled games
    are numerically unstable in the reference algorithm itself (indefinite proximity Hessians make
    the Riccati recursion grow to fp32 overflow -- the fp64 oracle reaches 1e72); where the oracle
    overflows, fp32 rounding decides which entries are inf/nan and there is nothing to compare."""
    r = np.asarray(ref, np.float64).reshape(len(ref), -1)
    return np.isfinite(r).all(axis=1) & (np.abs(r).max(axis=1, initial=0.0) < limit)


def close(a, b, tol=STAGE_TOL, what="", atol=1e-5, rows=None, cond=None):
    """max|a - b| <= tol * max|b| + atol per instance row (norm-wise; atol absorbs cancellation
    to ~0), over the rows where the oracle is tame.

    cond (optional): per-row max|o32 - o64| of the same quantity, the fp32 oracle's OWN distance to
    the fp64 answer.  A row is held to `tol` plus 4 x that distance: the tensor-core sweep does not
    share the oracle's rounding (3xTF32 products, different summation order), so on an
    ill-conditioned game -- one the oracle itself only resolves to 4e-4 -- it cannot sit within 1e-4
    of the fp32 oracle, only within a small multiple of the oracle's own error (measured:
    median 2.2 x for P, 1.0 x for alpha, tools/stage_errors.py).  Well-conditioned rows
    (cond ~ 1e-6 of scale) keep the plain tolerance.  Returns the number of rows compared."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape
    same_nonfinite = (a == b) & ~np.isfinite(b)           # e.g. merit = +inf on both sides
    a, b = np.where(same_nonfinite, 0.0, a), np.where(same_nonfinite, 0.0, b)
    finite = np.isfinite(b.reshape(len(b), -1)).all(axis=1)
    keep = finite if rows is None else (rows & finite)
    if not keep.any():
        return 0
    if a.ndim == 1:
        a, b = a[:, None], b[:, None]
    a, b = a[keep].reshape(keep.sum(), -1), b[keep].reshape(keep.sum(), -1)
    if a.size == 0:
        return 0
    scale = np.abs(b).max(axis=1)
    err = np.abs(a - b).max(axis=1)
    slack = 0.0 if cond is None else 4.0 * np.asarray(cond, np.float64)[keep]
    bad = ~(err <= tol * scale + atol + slack)
    assert not bad.any(), (f"{what}: worst row max|d| = {np.nanmax(err[bad]) if bad.any() else 0:.3e}, "
                           f"scale = {scale[bad][0]:.3e}, {bad.sum()} of {len(bad)} rows")
    return int(keep.sum())


def row_distance(o32, o64):
    """Per-row max|o32 - o64| (nan / inf -> inf): the `cond` argument of close()."""
    a = np.asarray(o32, np.float64).reshape(len(o32), -1)
    b = np.asarray(o64, np.float64).reshape(len(o64), -1)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b).max(axis=1)
    return np.where(np.isfinite(e), e, np.inf)


def wellposed(o32, o64, tol=1e-3):
    """Per-instance mask: the fp32 and fp64 builds of the oracle agree to `tol` (norm-wise).
    Where they do not, the instance is ill-conditioned or on a knife edge of the reference
    algorithm itself and fp32 implementations legitimately differ."""
    a = np.asarray(o32, np.float64).reshape(len(o32), -1)
    b = np.asarray(o64, np.float64).reshape(len(o64), -1)
    with np.errstate(invalid="ignore"):
        err = np.where(a == b, 0.0, np.abs(a - b)).max(axis=1)   # (inf == inf: the merit with the linesearch off)
        scale = np.where(np.isfinite(b), np.abs(b), 0.0).max(axis=1)
    return np.isfinite(err) & (err <= tol * scale + 1e-5)


def pair(product, oracle, name, batch, **param_overrides):
    build, params, x0f = CONFIGS[name]
    desc, _ = build()
    x0 = x0f(batch)
    hs = []
    for lib in (product, oracle):
        h = abi.Handle(lib, desc, params(**param_overrides), x0.shape[0], 0)
        h.upload_x0(x0)
        hs.append(h)
    return hs


# ------------------------------------------------------------------ index logic
@pytest.mark.parametrize("name", HEADLINE)
def test_layout_and_index_maps_bit_exact(product, oracle, name):
    c, o = pair(product, oracle, name, 2)
    for field in ("num_time_steps", "num_players", "xdim", "total_udim", "num_pairs", "R_floats",
                  "r_floats", "num_constraints"):
        assert getattr(c.layout, field) == getattr(o.layout, field), field
    for arr in ("udim", "u_offset", "pair_player", "pair_arg", "pair_R_offset", "pair_r_offset",
                "lambda_index"):
        assert list(getattr(c.layout, arr)) == list(getattr(o.layout, arr)), arr
    assert c.layout.record_floats % 4 == 0 and c.layout.record_floats > 0


# ------------------------------------------------------------------ LQ solve
@pytest.mark.parametrize("nominal", [0.0, 0.5])
def test_lq_backward_reference_test_system(product, oracle, nominal):
    # LQFeedbackSolverTest.MatchesLyapunovIterations (test/test_lq_solver.cpp:292-317) on the GPU
    desc = problems.lq_only(100, 2, [1, 1], cross_pairs=[(0, 1), (1, 0)])
    arrs, (A, B1, B2, Q, R) = lq_test_system(nominal)
    outs = []
    for lib in (product, oracle):
        h = abi.Handle(lib, desc, abi.SolverParams.defaults(), 1)
        h.upload_lq(**arrs)
        h.lq_backward()
        outs.append((h.download(abi.LQ_PS)[0], h.download(abi.LQ_ALPHAS)[0]))
        h.close()
    (Pc, ac), (Po, ao) = outs
    close(Pc, Po, what="P")
    close(ac, ao, tol=1e-5 if nominal else 1.0, what="alpha")
    one = lambda v: np.array([[v]])
    P1, P2 = lyapunov_iterations(A, B1, B2, Q[0], Q[1], one(R[0]), one(R[1]), one(R[2]), one(R[3]))
    assert np.abs(P1 - Pc[0, 0:1]).max() < 1e-4  # kSmallNumber, the reference's own tolerance
    assert np.abs(P2 - Pc[0, 1:2]).max() < 1e-4
    assert np.all(Pc[-1] == 0) and np.all(ac[-1] == 0)


def test_lq_backward_gershgorin_and_batch(product, oracle):
    # random LQ games where the adaptive regularisation fires (indefinite R) -- batch of 5
    rng = np.random.default_rng(5)
    B, T, n = 5, 30, 2
    desc = problems.lq_only(T, n, [1, 1], cross_pairs=[(0, 1), (1, 0)])
    A = np.tile(np.eye(n) + 0.1 * rng.normal(size=(B, T, n, n)), 1).astype(np.float32)
    Bs = (0.3 * rng.normal(size=(B, T, n, 2))).astype(np.float32)
    Qh = rng.normal(size=(B, T, 2, n, n))
    Q = (Qh @ np.swapaxes(Qh, -1, -2) * 0.2 + 0.1 * np.eye(n)).astype(np.float32)
    l = rng.normal(size=(B, T, 2, n)).astype(np.float32)
    R = rng.normal(size=(B, T, 4)).astype(np.float32) * 0.3  # includes negative R_ii
    r = rng.normal(size=(B, T, 4)).astype(np.float32)
    res = []
    for lib in (product, oracle):
        h = abi.Handle(lib, desc, abi.SolverParams.defaults(), B)
        h.upload_lq(A, Bs, Q, l, R, r)
        h.lq_backward()
        res.append((h.download(abi.LQ_PS), h.download(abi.LQ_ALPHAS), h.download(abi.DELTA_XS)))
        h.close()
    for (x, y, nm) in zip(res[0], res[1], ("P", "alpha", "dxs")):
        close(x, y, tol=2e-3, what=nm)  # looser: random indefinite systems are ill-conditioned


@pytest.mark.parametrize("n,udims", [(16, (1, 2, 3)), (16, (1, 1, 4)), (24, (3, 1, 2, 2)), (10, (2, 3)), (7, (8,))])
def test_lq_backward_any_shape_and_nonuniform_controls(product, oracle, n, udims):
    """an LQ-only handle whose (n, M, N) is one of the shape-keyed kernels' -- (16, 6, 3),
    (24, 8, 4) -- but whose players do NOT have M / N controls each must not reach the half-warp
    kernel (it hard-codes m = M / N); like every other shape in the envelope it takes the
    run-time-dimension sweep.  LQFeedbackSolver::Solve takes any dimensions
    (src/lq_feedback_solver.cpp:71-244)."""
    rng = np.random.default_rng(n * 31 + len(udims))
    B, T, N, M = 6, 20, len(udims), sum(udims)
    desc = problems.lq_only(T, n, list(udims))
    A = (np.eye(n) + 0.05 * rng.normal(size=(B, T, n, n))).astype(np.float32)
    Bs = (0.2 * rng.normal(size=(B, T, n, M))).astype(np.float32)
    Qh = rng.normal(size=(B, T, N, n, n))
    Q = (Qh @ np.swapaxes(Qh, -1, -2) * 0.1 + 0.5 * np.eye(n)).astype(np.float32)
    l = rng.normal(size=(B, T, N, n)).astype(np.float32)
    res = []
    for lib in (product, oracle):
        h = abi.Handle(lib, desc, abi.SolverParams.defaults(), B)
        lo = h.layout
        # R_ii = a well-conditioned SPD block per player, r_ii random (the only pairs of this descriptor)
        R = np.zeros((B, T, lo.R_floats), np.float32)
        r = np.zeros((B, T, lo.r_floats), np.float32)
        rr = np.random.default_rng(7)
        for pidx in range(lo.num_pairs):
            m = udims[lo.pair_arg[pidx]]
            Rh = rr.normal(size=(B, T, m, m))
            blk = Rh @ np.swapaxes(Rh, -1, -2) * 0.1 + np.eye(m)
            R[..., lo.pair_R_offset[pidx]:lo.pair_R_offset[pidx] + m * m] = blk.reshape(B, T, m * m)
            r[..., lo.pair_r_offset[pidx]:lo.pair_r_offset[pidx] + m] = rr.normal(size=(B, T, m))
        h.upload_lq(A, Bs, Q, l, R, r)
        h.lq_backward()
        res.append((h.download(abi.LQ_PS), h.download(abi.LQ_ALPHAS), h.download(abi.DELTA_XS),
                    h.download(abi.EXPECTED_DECREASE)))
        h.close()
    for (x, y, nm) in zip(res[0], res[1], ("P", "alpha", "dxs", "expected decrease")):
        compared = close(x, y, tol=1e-3, what=f"{nm} n={n} udims={udims}")
        assert compared == B


# ------------------------------------------------------------------ stage-by-stage parity
# BASELINE.json's configs name T = 150 for RoundaboutMerging and T = 50 for Air3D (the reference's
# own examples use 100): both horizons are run
STAGE_CASES = [(n, 100) for n in HEADLINE] + [("roundabout_merging", 150), ("air_3d", 50)]


@pytest.mark.parametrize("name,T", STAGE_CASES)
def test_stage_parity(product, oracle, oracle64, name, T, iterations=2, min_wellposed=None):
    build, params, x0f = CONFIGS[name]
    desc, _ = build(num_time_steps=T)
    x0 = x0f(16)
    B = x0.shape[0]
    hs = []
    for lib in (product, oracle, oracle64):
        h = abi.Handle(lib, desc, params(), B, 0)
        h.upload_x0(x0)
        h.solve_begin()
        hs.append(h)
    c, o, o64 = hs
    good = np.ones(B, bool)  # instances still well posed (fp32 and fp64 oracles agree)
    compared = []
    flips = 0

    def check(what, tol=None, label=""):
        # the first iteration sees identical inputs; later ones inherit ~1e-6 input differences
        # amplified by the step (roundabout: 0.75), and the expected decrease is a cancelling sum
        # that the CUDA path accumulates through the adjoint recursion (ilqg_backward.cuh)
        if tol is None:
            tol = 1e-3 if (what == abi.EXPECTED_DECREASE or label.startswith("it1")) else STAGE_TOL
        nonlocal good
        a, b, b64 = c.download(what), o.download(what), o64.download(what)
        good = good & wellposed(b, b64)
        # conditioning-aware: the LQ solution's rounding is not the oracle's (tensor cores), and every later
        # stage inherits it -- amplified by the feedback rollout on the ill-conditioned examples
        cond = None if label.startswith("prologue") or "record" in label else row_distance(b, b64)
        compared.append(close(a, b, tol=tol, rows=good, what=f"{label} field {what}", cond=cond))

    for what in (abi.XS, abi.US, abi.TOTAL_COSTS):
        check(what, label="prologue")
    assert np.array_equal(c.download(abi.TIME_OF_EXTREME)[good], o.download(abi.TIME_OF_EXTREME)[good])
    for it in range(iterations):
        for h in hs:
            h.linearize_quadraticize()
        for what in (abi.LIN_A, abi.LIN_B, abi.QUAD_Q, abi.QUAD_L, abi.QUAD_R, abi.QUAD_RGRAD):
            check(what, label=f"it{it} record")
        for h in hs:
            h.lq_backward()
        for what in (abi.LQ_PS, abi.LQ_ALPHAS, abi.DELTA_XS, abi.EXPECTED_DECREASE):
            check(what, label=f"it{it} LQ")
        for h in hs:
            h.linesearch()
        good = good & (o.download(abi.BACKTRACKS) == o64.download(abi.BACKTRACKS))
        # control flow first: where the fp32 and fp64 oracles agree the CUDA path takes the same Armijo
        # decisions, up to the odd knife-edge game (its merit differs from the oracle's by rounding);
        # such a game is counted, bounded, and leaves the value comparisons that follow
        same = np.ones(B, bool)
        for what in (abi.STATUS, abi.ITERS, abi.BACKTRACKS):
            same &= c.download(what) == o.download(what)
        flips += int((good & ~same).sum())
        assert flow_ok(same, good), f"it{it}: control flow differs on {(good & ~same).sum()} of {good.sum()} well-posed games"
        good = good & same
        for what in (abi.XS, abi.US, abi.PS, abi.ALPHAS, abi.MERIT, abi.STEP, abi.TOTAL_COSTS):
            check(what, label=f"it{it} linesearch")
        assert np.array_equal(c.download(abi.TIME_OF_EXTREME)[good], o.download(abi.TIME_OF_EXTREME)[good])
    report(f"stage_parity[{name},T={T}]", batch=B, wellposed_at_end=good.sum(), min_rows_compared=min(compared),
           max_rows_compared=max(compared), armijo_flips=flips)
    # measured on B200 (round 2): 13 (intersection) / 11 (roundabout) / 16 (air3d) of 16 games stay
    # well posed through two iterations at T = 100; RoundaboutMerging at T = 150 (regularisation 0,
    # a 50 % longer Riccati sweep) keeps 7 -- the fp32 and fp64 ORACLES part on the other nine
    floor = 5 if (name, T) == ("roundabout_merging", 150) else B // 2
    if min_wellposed is not None:
        floor = min_wellposed
    assert good.sum() >= floor, f"only {good.sum()} of {B} well-posed instances left"
    assert min(compared) >= floor
    for h in hs:
        h.close()


# ------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", HEADLINE)
def test_against_golden_fixture(product, name):
    """Golden iterate `it` = state after `it` iLQ iterations.  The CUDA library schedules
    linesearches asynchronously, so iterate `it` is obtained by a solve capped at `it` iterations
    (a golden RUNNING status then reads MAX_ITERS here)."""
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    build, params, _ = CONFIGS[name]
    desc, _ = build()
    iters = max(int(k.split("_")[1]) for k in g.files if k.startswith("merit_"))
    xs_tol = 2e-2 if name == "roundabout_merging" else 1e-3  # roundabout: reg = 0, ill-conditioned
    total_compared = 0
    for it in range(1, iters + 1):
        h = abi.Handle(product, desc, params(max_solver_iters=it), g["x0"].shape[0], 0)
        h.upload_x0(g["x0"])
        h.solve_begin()
        if it == 1:
            close(h.download(abi.XS), g["xs_0"], what="xs_0")
            close(h.download(abi.TOTAL_COSTS), g["costs_0"], what="costs_0")
        h.solve(chunk=it)
        gstat = g[f"status_{it}"].copy()
        gstat[gstat == abi.STATUS_RUNNING] = abi.STATUS_MAX_ITERS
        flow = (h.download(abi.BACKTRACKS) == g[f"backtracks_{it}"]) & (h.download(abi.STATUS) == gstat) & (
            h.download(abi.ITERS) == g[f"iters_{it}"])
        stable = g[f"stable_{it}"]
        # on instances where the fp32 and fp64 oracles agree the CUDA path must follow the same flow
        ok = flow & stable & (g[f"status_{it}"] != abi.STATUS_LINESEARCH_FAILED) & tame(g[f"xs_{it}"], 1e3)
        report(f"golden_fixture[{name}] it{it}", batch=len(flow), stable=stable.sum(), flow_on_stable=flow[stable].mean(),
               rows_compared=ok.sum())
        assert flow_ok(flow, stable), f"iteration {it}: flow matches on {flow[stable].mean():.0%} of stable"
        assert ok.sum() >= max(2, int(0.5 * stable.sum())), f"iteration {it}: only {ok.sum()} rows comparable"
        n_cmp = close(h.download(abi.XS), g[f"xs_{it}"], tol=xs_tol, atol=1e-3, rows=ok, what=f"xs_{it}")
        close(h.download(abi.US), g[f"us_{it}"], tol=xs_tol, atol=1e-3, rows=ok, what=f"us_{it}")
        assert np.array_equal(h.download(abi.TIME_OF_EXTREME)[ok], g[f"t_extreme_{it}"][ok])
        close(h.download(abi.MERIT), g[f"merit_{it}"], tol=max(1e-3, xs_tol), rows=ok, what="merit")
        total_compared += n_cmp
        h.close()
    assert g[f"stable_{iters}"].sum() >= 3, "golden fixture has too few well-posed instances"
    assert total_compared >= 2 * iters


# ------------------------------------------------------------------ reference fixtures
@pytest.mark.parametrize("name", HEADLINE)
def test_against_reference_fixture(product, oracle64, name, min_compared=6):
    """The CUDA path against outputs of the reference's own sources (tests/golden/ref_*.npz, made
    by tests/golden/make_ref_golden.py; the CPU oracle reproduces them bit for bit in
    tests/test_ref_pins.py).  Log iterate `it` of ILQSolver::Solve = the operating point after
    `it` iterations; instances are compared where the fp64 build of the oracle agrees with the
    reference's fp32 result to 1e-3 (elsewhere fp32 rounding decides the Armijo branch)."""
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    build, params, _ = CONFIGS[name]
    desc, _ = build()
    x0 = g["x0"]
    B = x0.shape[0]
    iters = int(g["ilq_iters"])
    xs_tol = 2e-2 if name == "roundabout_merging" else 1e-3
    compared = stable_total = 0
    for it in range(0, iters + 1):
        hs = []
        for lib in (product, oracle64):
            h = abi.Handle(lib, desc, params(max_solver_iters=max(it, 1)), B, 0)
            h.upload_x0(x0)
            h.solve_begin()
            if it > 0:
                h.solve(chunk=it)
            hs.append(h)
        c, o64 = hs
        logged = g["ilq_iterates"] > it           # the reference logged iterate `it`
        ref_xs, ref_us = g["ilq_xs"][:, it], g["ilq_us"][:, it]
        if it == 0:
            close(c.download(abi.XS), ref_xs, what="initial rollout xs")
            close(c.download(abi.US), ref_us, what="initial rollout us")
            continue
        stable = logged & tame(np.where(np.isfinite(ref_xs), ref_xs, 0.0), 1e3)
        if stable.any():
            stable[stable] &= wellposed(ref_xs[stable], o64.download(abi.XS)[stable])
        stable &= (o64.download(abi.ITERS) == it) & (o64.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED)
        flow = (c.download(abi.ITERS) == it) & (c.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED)
        ok = stable & flow
        report(f"reference_fixture[{name}] it{it}", batch=B, stable=stable.sum(),
               flow_on_stable=flow[stable].mean() if stable.any() else -1.0, rows_compared=ok.sum())
        if stable.any():
            assert flow_ok(flow, stable), f"iterate {it}: CUDA path follows the reference on {flow[stable].mean():.0%}"
        # (conditioning-aware like the stage tests: 4 x the reference's own distance to the fp64 answer)
        close(c.download(abi.XS), ref_xs, tol=xs_tol, atol=1e-3, rows=ok, what=f"ref xs_{it}",
              cond=row_distance(np.where(np.isfinite(ref_xs), ref_xs, 0.0), o64.download(abi.XS)))
        close(c.download(abi.US), ref_us, tol=xs_tol, atol=1e-3, rows=ok, what=f"ref us_{it}",
              cond=row_distance(np.where(np.isfinite(ref_us), ref_us, 0.0), o64.download(abi.US)))
        compared += int(ok.sum())
        stable_total += int(stable.sum())
        for h in hs:
            h.close()
    assert compared >= max(min_compared, int(0.8 * stable_total)), f"only {compared} of {stable_total} stable (instance, iterate) pairs compared"


# ------------------------------------------------------------------ open-loop LQ solver
def test_open_loop_lq_reference_test_system(product, oracle):
    """LQOpenLoopSolver::Solve on the system of LQOpenLoopSolverTest (test/test_lq_solver.cpp:
    143-177, 347-379; x0 = ones, nominal 0.5): CUDA vs oracle, alphas and delta_xs."""
    from tests.test_oracle_pins import solve_lq_open_loop
    x0 = np.ones(2)
    Pc, ac, dc, _ = solve_lq_open_loop(product, 0.5, x0)
    Po, ao, do, _ = solve_lq_open_loop(oracle, 0.5, x0)
    assert np.all(Pc == 0) and np.all(Po == 0)
    close(ac[None], ao[None], tol=1e-4, what="open-loop alphas")
    close(dc[None], do[None], tol=1e-4, what="open-loop delta_xs")


@pytest.mark.parametrize("name", HEADLINE)
def test_open_loop_stage_parity(product, oracle, oracle64, name):
    """One LQOpenLoopSolver::Solve on the LQ records of the initial rollout of each example:
    alphas, delta_xs, expected decrease against the oracle, on the instances where the fp32 and
    fp64 oracles agree."""
    build, params, x0f = CONFIGS[name]
    desc, _ = build()
    x0 = x0f(16)
    hs = []
    for lib in (product, oracle, oracle64):
        h = abi.Handle(lib, desc, params(open_loop=1), x0.shape[0], 0)
        h.upload_x0(x0)
        h.solve_begin()
        h.linearize_quadraticize()
        h.lq_backward()
        hs.append(h)
    c, o, o64 = hs
    good = wellposed(o.download(abi.LQ_ALPHAS), o64.download(abi.LQ_ALPHAS)) & tame(o.download(abi.LQ_ALPHAS))
    assert good.sum() >= 4, f"only {good.sum()} well-posed instances"
    assert np.all(c.download(abi.LQ_PS) == 0)
    close(c.download(abi.LQ_ALPHAS), o.download(abi.LQ_ALPHAS), tol=1e-3, rows=good, what="alphas")
    close(c.download(abi.DELTA_XS), o.download(abi.DELTA_XS), tol=1e-3, rows=good, what="delta_xs")
    close(c.download(abi.EXPECTED_DECREASE), o.download(abi.EXPECTED_DECREASE), tol=1e-3, atol=1e-3,
          rows=good, what="expected decrease")
    for h in hs:
        h.close()


@pytest.mark.parametrize("name", ["three_player_intersection", "roundabout_merging"])
def test_open_loop_against_reference_fixture(product, oracle64, name):
    """ILQSolver on LQOpenLoopSolver: the CUDA path against the iterates the reference's own sources
    logged (tests/golden/ref_*.npz, keys ol_*), on the instances where the fp64 build of the oracle
    agrees with the reference's fp32 result (elsewhere rounding decides the Armijo branch)."""
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    build, params, _ = CONFIGS[name]
    desc, _ = build()
    nol = g["ol_xs"].shape[0]
    iters = int(g["ol_iters"])
    compared = 0
    for it in range(1, iters + 1):
        hs = []
        for lib in (product, oracle64):
            h = abi.Handle(lib, desc, params(max_solver_iters=it, open_loop=1), nol, 0)
            h.upload_x0(g["x0"][:nol])
            h.solve_begin()
            h.solve(chunk=it)
            hs.append(h)
        c, o64 = hs
        ref_xs, ref_us = g["ol_xs"][:, it], g["ol_us"][:, it]
        stable = g["ol_iterates"] > it
        if stable.any():
            stable[stable] &= wellposed(ref_xs[stable], o64.download(abi.XS)[stable])
        stable &= (o64.download(abi.ITERS) == it) & (o64.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED)
        flow = (c.download(abi.ITERS) == it) & (c.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED)
        ok = stable & flow
        report(f"open_loop_fixture[{name}] it{it}", batch=nol, stable=stable.sum(),
               flow_on_stable=flow[stable].mean() if stable.any() else -1.0, rows_compared=ok.sum())
        assert flow_ok(flow, stable)
        close(c.download(abi.XS), ref_xs, tol=2e-3, atol=1e-3, rows=ok, what=f"open-loop xs_{it}")
        close(c.download(abi.US), ref_us, tol=2e-3, atol=1e-3, rows=ok, what=f"open-loop us_{it}")
        compared += int(ok.sum())
        for h in hs:
            h.close()
    assert compared >= 6, f"only {compared} (instance, iterate) pairs were comparable"


# ------------------------------------------------------------------ receding horizon
def test_receding_horizon_against_reference_fixture(product):
    """Problem::SetUpNextRecedingHorizon on the device (k_receding_horizon) against what the
    reference's own sources produce (tests/golden/ref_roundabout_merging.npz, keys rh_*): new t0
    exact, states / controls / strategies of the re-based plan within fp32 tolerance."""
    from tests.test_ref_pins import receding_horizon_cases
    g = np.load(os.path.join(GOLDEN, "ref_roundabout_merging.npz"))
    build, params, _ = CONFIGS["roundabout_merging"]
    desc, _ = build()
    # from the reference's plan, so that the comparison is about this step and not about the
    # (ill-conditioned, reg = 0) RoundaboutMerging solve before it
    for c, h, new_t0 in receding_horizon_cases(product, g, desc, params, from_plan=True):
        assert new_t0 == g["rh_t0"][c]
        close(h.download(abi.X0), g["rh_x0"][c], tol=1e-3, what=f"case {c} x0")  # one plan has |x| ~ 2e3
        close(h.download(abi.WARM_XS), g["rh_xs"][c], tol=1e-3, atol=1e-3, what=f"case {c} xs")
        close(h.download(abi.WARM_US), g["rh_us"][c], tol=1e-3, atol=1e-3, what=f"case {c} us")
        close(h.download(abi.WARM_ALPHAS), g["rh_alphas"][c], tol=1e-3, atol=1e-3, what=f"case {c} alphas")
        close(h.download(abi.WARM_PS), g["rh_Ps"][c], tol=1e-3, atol=1e-3, what=f"case {c} Ps")
        # the zero extension is exact
        ref_zero = g["rh_Ps"][c] == 0
        assert np.all(h.download(abi.WARM_PS)[ref_zero.all(axis=(2, 3))] == 0)


def wellposed_rows(in_double, want, tol=1e-5):
    """Games where the oracle's fp64 build agrees with the reference's fp32 result: elsewhere (the
    reg = 0 RoundaboutMerging plans carry gains ~5e3) fp32 rounding alone moves the answer."""
    a, b = np.asarray(in_double, np.float64), np.asarray(want, np.float64)
    a, b = a.reshape(len(a), -1), b.reshape(len(b), -1)
    return np.abs(a - b).max(axis=1) <= tol * np.maximum(1.0, np.abs(b).max(axis=1))


@pytest.mark.parametrize("name", ["three_player_intersection", "roundabout_merging", "air_3d"])
def test_integrate_plan_against_reference_fixture(product, oracle64, name):
    """MultiPlayerIntegrableSystem::Integrate(t0, t, x0, plan) on the device (k_integrate_plan)
    against what the reference's own sources return (tests/golden/ref_integrate_<name>.npz): plans
    of the horizon's length and five steps longer, t0 at / past the plan's start, both times inside
    one step, t == t0, grid times.  fp32 tolerance (the device contracts a*b+c into fma), on the
    games where the computation is well posed."""
    from tests.test_ref_pins import integrate_plan_cases
    n_cases = n_rows = 0
    for (tag, c, got, want), (_, _, in_double, _) in zip(integrate_plan_cases(product, name),
                                                        integrate_plan_cases(oracle64, name)):
        assert np.isfinite(got).all()
        rows = wellposed_rows(in_double, want)
        close(got, want, tol=1e-4, atol=1e-4, what=f"{tag} case {c}", rows=rows)
        n_cases += 1
        n_rows += int(rows.sum())
    assert n_cases == 9 and n_rows >= 18


def test_receding_horizon_from_spliced_plan_against_reference_fixture(product, oracle64):
    """OverwriteSolution(spliced plan, 105 steps) + SetUpNextRecedingHorizon as the receding-horizon
    simulator runs them (src/receding_horizon_simulator.cpp:105-109), keys rhl_* of
    tests/golden/ref_integrate_roundabout_merging.npz."""
    from tests.test_ref_pins import long_plan_receding_cases
    T = 100
    compared = 0
    for (c, h, new_t0, gi), (_, h64, _, _) in zip(long_plan_receding_cases(product), long_plan_receding_cases(oracle64)):
        assert new_t0 == gi["rhl_t0"][c]
        rows = wellposed_rows(h64.download(abi.WARM_XS)[:, :T], gi["rhl_xs"][c]) & wellposed_rows(
            h64.download(abi.X0), gi["rhl_x0"][c])
        close(h.download(abi.X0), gi["rhl_x0"][c], tol=1e-3, what=f"case {c} x0", rows=rows)
        for what, key in ((abi.WARM_XS, "rhl_xs"), (abi.WARM_US, "rhl_us"), (abi.WARM_PS, "rhl_Ps"),
                          (abi.WARM_ALPHAS, "rhl_alphas")):
            close(h.download(what)[:, :T], gi[key][c], tol=1e-3, atol=1e-3, what=f"case {c} {key}", rows=rows)
        compared += int(rows.sum())
    assert compared >= 4


# ------------------------------------------------------------------ full solves
# measured round 2 (B200): see profiles/r02_parity_counts.md; thresholds = measured minus a margin
# (intersection 0.92, roundabout 0.75, air3d 0.97 of the whole batch, ill-posed games included)
FULL_SOLVE_FLOW_MIN = {"three_player_intersection": 0.85, "roundabout_merging": 0.65, "air_3d": 0.9}


@pytest.mark.parametrize("name,batch,iters", [("three_player_intersection", 64, 10),
                                              ("roundabout_merging", 32, 3), ("air_3d", 36, 10)])
def test_full_solve_against_oracle(product, oracle, name, batch, iters):
    c, o = pair(product, oracle, name, batch, max_solver_iters=iters)
    for h in (c, o):
        h.solve_begin()
        h.solve(chunk=iters)
    flow = (c.download(abi.STATUS) == o.download(abi.STATUS)) & (
        c.download(abi.ITERS) == o.download(abi.ITERS)) & (
        c.download(abi.BACKTRACKS) == o.download(abi.BACKTRACKS))
    ok = flow & (o.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED) & tame(o.download(abi.XS), 1e3)
    solvable = (o.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED) & tame(o.download(abi.XS), 1e3)
    report(f"full_solve[{name}]", batch=len(flow), flow=flow.mean(), oracle_solvable=solvable.sum(), rows_compared=ok.sum())
    # control flow over `iters` chained linesearches: every fp32 implementation parts from another on
    # the knife-edge games (the oracle's own fp32 and fp64 builds do), hence a fraction, not equality
    assert flow.mean() >= FULL_SOLVE_FLOW_MIN[name], f"control flow identical for only {flow.mean():.0%}"
    assert ok.sum() >= max(2, int(0.5 * solvable.sum())), f"only {ok.sum()} of {solvable.sum()} solvable games compared"
    xs_c, xs_o = c.download(abi.XS)[ok], o.download(abi.XS)[ok]
    assert np.isfinite(xs_c).all()
    rel = np.abs(xs_c - xs_o).max(axis=(1, 2)) / np.maximum(1.0, np.abs(xs_o).max(axis=(1, 2)))
    assert np.median(rel) < 1e-3


# ------------------------------------------------------------------ AL outer update
def test_augmented_lagrangian_update(product, oracle):
    c, o = pair(product, oracle, "three_player_intersection", 8, max_solver_iters=3)
    for h in (c, o):
        h.solve_begin()
        h.solve(chunk=3)
        h.al_update()
    close(c.download(abi.LAMBDAS), o.download(abi.LAMBDAS), what="lambdas")
    close(c.download(abi.MU), o.download(abi.MU), what="mu")
    close(c.download(abi.MAX_CONSTRAINT_ERROR), o.download(abi.MAX_CONSTRAINT_ERROR), what="max err")
    lam = c.download(abi.LAMBDAS)
    # lambda slots 43, 81, 86, 91 are never touched by the sweep
    assert np.all(lam[:, :, [43, 81, 86, 91]] == 0.0)
    for h in (c, o):
        h.overwrite_solution(only_successful=True)
        h.solve_begin()
        h.solve(chunk=3)
        h.al_post_solve()
    assert np.array_equal(c.download(abi.STATUS), o.download(abi.STATUS))
    close(c.download(abi.XS), o.download(abi.XS), tol=1e-3, what="xs after AL step")
    close(c.download(abi.MU), o.download(abi.MU), what="mu after post-solve")


def test_augmented_lagrangian_solve(product, oracle, oracle64):
    """Row f1: the whole AugmentedLagrangianSolver::Solve per game (ilqg_al_begin / ilqg_al_advance).
    An AL solve chains ~10 inner solves, so only games whose fp32 and fp64 oracle runs agree
    (same iterate count, same success, close trajectory) are compared by value."""
    from ilqgames_b200 import al
    B, cap, tol = 12, 40, 0.1
    desc, _ = problems.three_player_intersection()
    x0 = problems.three_player_intersection_x0_batch(B, 2024)
    res = {}
    for name, lib in (("cuda", product), ("o32", oracle), ("o64", oracle64)):
        params = problems.three_player_intersection_params()
        params.max_solver_iters = params.unconstrained_solver_max_iters
        h = abi.Handle(lib, desc, params, B, 0)
        h.upload_x0(x0)
        res[name] = al.solve_augmented_lagrangian(h, cap, tol, reset_problem=False, reset_lambdas=False,
                                                  reset_mu=False)
        assert (h.download(abi.AL_STATE) == 2).all()
        res[name + "_mu"] = h.download(abi.MU)
    c, o, o64 = res["cuda"], res["o32"], res["o64"]
    # invariants of the loop itself, every game
    assert (c.iterates >= 1).all() and (c.iterates < cap + 11).all()
    assert not (c.success.astype(bool) & (c.max_constraint_error > tol)).any()
    assert c.rounds >= 2

    def agree(a, b):
        same = (a.iterates == b.iterates) & (a.success == b.success)
        err = np.abs(a.xs.astype(np.float64) - b.xs).reshape(B, -1).max(axis=1)
        scale = np.abs(b.xs).reshape(B, -1).max(axis=1)
        return same & (err <= 2e-2 * scale + 1e-3)

    stable = agree(o, o64)
    assert stable.sum() >= 3, f"only {stable.sum()} of {B} games are well-posed in fp32"
    match = agree(c, o)
    assert match[stable].mean() >= 0.75, (match, stable, c.iterates, o.iterates)
    close(res["cuda_mu"][stable & match], res["o32_mu"][stable & match], tol=1e-4, what="mu after the AL solve")


# ------------------------------------------------------------------ full-size properties
def test_full_batch_properties(product):
    """BASELINE.json's metric size (batch 4096, T = 100): properties that need no oracle run."""
    B, iters = 4096, 3
    desc, _ = problems.three_player_intersection()
    params = problems.three_player_intersection_params(max_solver_iters=iters, disable_convergence_exit=1)
    x0 = problems.three_player_intersection_x0_batch(B, 4096)
    h = abi.Handle(product, desc, params, B, 0)
    h.upload_x0(x0)
    h.solve_begin()
    h.solve(chunk=iters)
    status, its, rolls = h.download(abi.STATUS), h.download(abi.ITERS), h.download(abi.BACKTRACKS)
    xs, us, Ps, al = (h.download(w) for w in (abi.XS, abi.US, abi.PS, abi.ALPHAS))
    assert set(np.unique(status)) <= {abi.STATUS_MAX_ITERS, abi.STATUS_LINESEARCH_FAILED}
    assert np.all(its[status == abi.STATUS_MAX_ITERS] == iters)
    assert np.all(rolls >= its)
    np.testing.assert_array_equal(xs[:, 0], x0)           # every trajectory starts at its x0
    assert np.all(Ps[:, -1] == 0) and np.all(al[:, -1] == 0)
    ok = status == abi.STATUS_MAX_ITERS
    assert ok.mean() > 0.8 and np.isfinite(xs[ok]).all() and np.isfinite(us[ok]).all()
    # batch-position invariance: instances are independent, so a sub-batch solved alone must
    # reproduce its rows of the big batch bit for bit
    idx = np.array([0, 1, 777, 2048, 4095])
    h2 = abi.Handle(product, desc, params, len(idx), 0)
    h2.upload_x0(x0[idx])
    h2.solve_begin()
    h2.solve(chunk=iters)
    np.testing.assert_array_equal(h2.download(abi.XS), xs[idx])
    np.testing.assert_array_equal(h2.download(abi.US), us[idx])
    np.testing.assert_array_equal(h2.download(abi.BACKTRACKS), rolls[idx])
    # determinism: a second solve from the same warm start and solver state repeats the first
    h.reset(h.RESET_SOLVER)
    h.solve_begin()
    h.solve(chunk=iters)
    np.testing.assert_array_equal(h.download(abi.XS), xs)
    # the rollout is consistent with the dynamics: x_{k+1} = RK4(x_k, u_k) re-evaluated in numpy
    def f(x, u):
        th1, ph1, v1, a1 = x[:, 2], x[:, 3], x[:, 4], x[:, 5]
        th2, ph2, v2, a2 = x[:, 8], x[:, 9], x[:, 10], x[:, 11]
        th3, v3 = x[:, 14], x[:, 15]
        z = np.zeros_like(x)
        z[:, 0], z[:, 1], z[:, 2], z[:, 3], z[:, 4], z[:, 5] = (v1 * np.cos(th1), v1 * np.sin(th1),
                                                              v1 / 4.0 * np.tan(ph1), u[:, 0], a1, u[:, 1])
        z[:, 6], z[:, 7], z[:, 8], z[:, 9], z[:, 10], z[:, 11] = (v2 * np.cos(th2), v2 * np.sin(th2),
                                                                 v2 / 4.0 * np.tan(ph2), u[:, 2], a2, u[:, 3])
        z[:, 12], z[:, 13], z[:, 14], z[:, 15] = v3 * np.cos(th3), v3 * np.sin(th3), u[:, 4], u[:, 5]
        return z
    sel = np.nonzero(ok & tame(xs, 100.0))[0][:256]  # wild trajectories amplify fp32 rounding
    for k in (0, 37, 98):
        x, u = xs[sel, k].astype(np.float64), us[sel, k].astype(np.float64)
        for _ in range(2):
            k1 = 0.05 * f(x, u)
            k2 = 0.05 * f(x + 0.5 * k1, u)
            k3 = 0.05 * f(x + 0.5 * k2, u)
            k4 = 0.05 * f(x + k3, u)
            x = x + (k1 + 2 * (k2 + k3) + k4) / 6.0
        ref = xs[sel, k + 1].astype(np.float64)
        assert (np.abs(x - ref) / np.maximum(1.0, np.abs(ref))).max() < 1e-4
    h.close()
    h2.close()


def test_full_batch_sampled_rows_against_oracle(product, oracle, oracle64):
    """BASELINE.json's metric size against the oracle itself: the batch-4096 solve of the bench
    workload (seed 4096), 256 rows sampled across the batch re-solved by the fp32 and fp64 oracles.
    First iteration by value on every well-posed row; after three iterations control flow and
    trajectories on the rows where the two oracles agree."""
    B, iters = 4096, 3
    desc, _ = problems.three_player_intersection()
    x0 = problems.three_player_intersection_x0_batch(B, 4096)
    idx = np.sort(np.random.default_rng(7).choice(B, 256, replace=False))
    for it, tol in ((1, 1e-3), (iters, 2e-3)):
        params = problems.three_player_intersection_params(max_solver_iters=it, disable_convergence_exit=1)
        res = []
        for lib, x in ((product, x0), (oracle, x0[idx]), (oracle64, x0[idx])):
            h = abi.Handle(lib, desc, params, len(x), 0)
            h.upload_x0(x)
            h.solve_begin()
            h.solve(chunk=it)
            sel = idx if len(x) == B else slice(None)
            res.append({w: h.download(w)[sel] for w in (abi.XS, abi.US, abi.STATUS, abi.ITERS, abi.BACKTRACKS)})
            h.close()
        c, o, o64 = res
        same = lambda a, b: (a[abi.STATUS] == b[abi.STATUS]) & (a[abi.ITERS] == b[abi.ITERS]) & (
            a[abi.BACKTRACKS] == b[abi.BACKTRACKS])
        stable = same(o, o64) & wellposed(o[abi.XS], o64[abi.XS]) & tame(o[abi.XS], 1e3) & (
            o[abi.STATUS] != abi.STATUS_LINESEARCH_FAILED)
        flow = same(c, o)
        ok = stable & flow
        report(f"full_batch_sampled it{it}", sampled=len(idx), stable=stable.sum(), flow_on_stable=flow[stable].mean(),
               rows_compared=ok.sum())
        assert stable.sum() >= 192   # measured: 236 / 226 of 256
        assert flow_ok(flow, stable)
        close(c[abi.XS], o[abi.XS], tol=tol, atol=1e-3, rows=ok, what=f"xs after {it} iterations")
        close(c[abi.US], o[abi.US], tol=tol, atol=1e-3, rows=ok, what=f"us after {it} iterations")


# ------------------------------------------------------------------ ragged sizes
@pytest.mark.parametrize("batch,T", [(1, 100), (33, 100), (100, 23), (7, 2), (65, 57)])
def test_ragged_batch_and_horizon(product, oracle, oracle64, batch, T):
    """Batches that do not fill a warp / block and horizons that are not a multiple of the merit
    kernel's time-step chunk (down to T = 2, the smallest a Problem can have): same control flow
    and trajectories as the oracle after two iterations (on the instances where the fp32 and fp64
    oracles agree)."""
    desc, _ = problems.three_player_intersection(num_time_steps=T)
    params = problems.three_player_intersection_params(max_solver_iters=2)
    x0 = problems.three_player_intersection_x0_batch(batch, 31 + batch)
    hs = []
    for lib in (product, oracle, oracle64):
        h = abi.Handle(lib, desc, params, batch, 0)
        h.upload_x0(x0)
        h.solve_begin()
        hs.append(h)
    c, o, o64 = hs
    close(c.download(abi.XS), o.download(abi.XS), what="initial rollout")
    close(c.download(abi.TOTAL_COSTS), o.download(abi.TOTAL_COSTS), tol=1e-4, what="initial costs")
    for h in hs:
        h.solve(chunk=2)
    flow = (c.download(abi.STATUS) == o.download(abi.STATUS)) & (c.download(abi.ITERS) == o.download(abi.ITERS)) & (
        c.download(abi.BACKTRACKS) == o.download(abi.BACKTRACKS))
    ok = flow & (o.download(abi.STATUS) != abi.STATUS_LINESEARCH_FAILED) & tame(o.download(abi.XS), 1e3)
    ok &= wellposed(o.download(abi.XS), o64.download(abi.XS))
    report(f"ragged[{batch},{T}]", batch=batch, flow=flow.mean(), rows_compared=ok.sum())
    assert flow.mean() >= 0.7, f"control flow identical for only {flow.mean():.0%}"
    assert ok.sum() >= max(1, batch // 2), f"only {ok.sum()} of {batch} rows comparable"
    close(c.download(abi.XS), o.download(abi.XS), tol=2e-3, atol=1e-3, rows=ok, what="xs after 2 iterations")
    close(c.download(abi.US), o.download(abi.US), tol=2e-3, atol=1e-3, rows=ok, what="us after 2 iterations")
    np.testing.assert_array_equal(c.download(abi.XS)[:, 0], x0)
    for h in hs:
        h.close()


# ------------------------------------------------------------------ multi-GPU behind the API
def test_multi_device_handle_equals_single_device(product):
    """ilqg_create_multi (SURVEY 8e behind the product API, not only under torchrun): the batch sharded
    over two GPUs of the box by ONE handle reproduces the single-GPU solve bit for bit (the games are
    independent and every device runs the same kernels).  Skipped on a one-GPU box."""
    desc, _ = problems.three_player_intersection()
    params = problems.three_player_intersection_params(max_solver_iters=3, disable_convergence_exit=1)
    B = 257  # odd on purpose: the shards differ by one game
    x0 = problems.three_player_intersection_x0_batch(B, 11)
    try:
        hm = abi.Handle(product, desc, params, B, [0, 1])
    except abi.IlqgError as e:
        pytest.skip(f"needs two GPUs: {e}")
    outs = []
    for h in (abi.Handle(product, desc, params, B, 0), hm):
        h.upload_x0(x0)
        h.solve_begin()
        h.solve(chunk=3)
        outs.append([h.download(w) for w in (abi.XS, abi.US, abi.PS, abi.ALPHAS, abi.STATUS, abi.ITERS, abi.BACKTRACKS,
                                             abi.LIN_A, abi.QUAD_Q)])
        h.close()
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)


def test_standalone_lq_solve_with_nonzero_x0(product, oracle, oracle64):
    """LQFeedbackSolver::Solve(lin, quad, x0 != 0) on a (16, 6, 3) handle: delta_xs start at
    x0 and ExpectedDecrease includes the terms x0 drives -- the tensor-core sweep adds them to its
    adjoint sum, the dense-record path takes the forward sweep."""
    desc, _ = problems.three_player_intersection()
    x0 = problems.three_player_intersection_x0_batch(8, 5)
    lq_x0 = (0.05 * np.random.default_rng(3).normal(size=x0.shape)).astype(np.float32)
    res = []
    for lib in (product, oracle, oracle64):
        h = abi.Handle(lib, desc, problems.three_player_intersection_params(), 8, 0)
        h.upload_x0(x0)
        h.solve_begin()
        h.linearize_quadraticize()
        h.upload(abi.LQ_X0, lq_x0)
        h.lq_backward()
        res.append((h.download(abi.DELTA_XS), h.download(abi.EXPECTED_DECREASE)))
        h.close()
    (dc, ec), (do, eo), (d64, e64) = res
    good = wellposed(do, d64) & wellposed(eo, e64)
    assert good.sum() >= 4
    np.testing.assert_array_equal(dc[:, 0], lq_x0)
    close(dc, do, tol=1e-3, rows=good, what="delta_xs", cond=row_distance(do, d64))
    close(ec, eo, tol=1e-3, atol=1e-3, rows=good, what="expected decrease", cond=row_distance(eo, e64))


# ------------------------------------------------------------------ the two rollout kernels
ROLLOUT_CASES = [("three_player_intersection", 100, 67, 4), ("roundabout_merging", 150, 21, 3), ("air_3d", 50, 64, 4),
                 ("three_player_overtaking", 100, 9, 2), ("two_player_reachability", 100, 8, 2),
                 ("dubins_origin", 100, 8, 2), ("modified_air_3d", 100, 8, 2),
                 ("two_player_collision_avoidance_reachability", 100, 8, 2)]


@pytest.mark.parametrize("name,T,batch,iters", ROLLOUT_CASES)
def test_rollout_kernels_bit_identical(product, name, T, batch, iters):
    """The stage-parallel rollout (k_ls_rollout_sp: eight lanes per (item, subsystem), one per RK4
    stage) performs the floating-point operations of the lane-per-item rollout (k_ls_rollout,
    ILQG_ROLLOUT=lanes) on the same operands: whole solves -- every trajectory, strategy, merit,
    Armijo decision and counter -- are bit-identical, for every subsystem kind."""
    build, params, x0f = CONFIGS[name]
    desc, _ = build(num_time_steps=T) if name in HEADLINE + ["three_player_overtaking"] else build()
    x0 = x0f(batch)
    outs = []
    for kernel in ("lanes", "sp"):
        os.environ["ILQG_ROLLOUT"] = kernel
        try:
            h = abi.Handle(product, desc, params(max_solver_iters=iters, disable_convergence_exit=1), x0.shape[0], 0)
        finally:
            del os.environ["ILQG_ROLLOUT"]
        h.upload_x0(x0)
        h.solve_begin()
        first = [h.download(w) for w in (abi.XS, abi.US, abi.MERIT)]
        h.solve(chunk=iters)
        outs.append(first + [h.download(w) for w in (abi.XS, abi.US, abi.PS, abi.ALPHAS, abi.MERIT, abi.STATUS, abi.ITERS,
                                                    abi.BACKTRACKS, abi.TOTAL_COSTS, abi.STEP)])
        h.close()
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)
    report(f"rollout_bit_identical[{name},T={T}]", batch=x0.shape[0], iterations=iters,
           backtracks=int(outs[0][-3].sum()))


@pytest.mark.parametrize("name,T,batch", [("three_player_intersection", 100, 96), ("air_3d", 50, 128), ("roundabout_merging", 100, 24)])
def test_tier_splits_and_pipeline_give_identical_solves(product, name, T, batch):
    """How the continued linesearch is cut into tiers (ILQG_LS_TIERS), which rollout kernel a window
    takes (ILQG_ROLLOUT) and whether the open linesearches finish on a side stream (ILQG_PIPELINE) are
    schedules, not algorithms: candidate j's trajectory and merit depend on j alone and the first
    candidate that passes Armijo wins, so every variant must reproduce the one-window, one-stream
    solve bit for bit -- trajectories, strategies, merits, statuses, iteration and rollout counts."""
    build, params, x0f = CONFIGS[name]
    desc, _ = build(num_time_steps=T)
    x0 = x0f(batch)
    variants = [{"ILQG_LS_TIERS": "99", "ILQG_PIPELINE": "0", "ILQG_ROLLOUT": "sp"},
                {}, {"ILQG_LS_TIERS": "7,32"}, {"ILQG_LS_TIERS": "5"}, {"ILQG_LS_TIERS": "39", "ILQG_PIPELINE": "0"},
                {"ILQG_ROLLOUT": "lanes"}, {"ILQG_LS_TIERS": "7,32", "ILQG_PIPELINE": "1", "ILQG_ROLLOUT": "auto"}]
    outs = []
    for env in variants:
        os.environ.update(env)
        try:
            h = abi.Handle(product, desc, params(max_solver_iters=6, disable_convergence_exit=1), x0.shape[0], 0)
        finally:
            for k in env:
                del os.environ[k]
        h.upload_x0(x0)
        h.solve_begin()
        h.solve(chunk=6)
        outs.append([h.download(w) for w in (abi.XS, abi.US, abi.PS, abi.ALPHAS, abi.MERIT, abi.STATUS, abi.ITERS,
                                             abi.BACKTRACKS, abi.TOTAL_COSTS, abi.STEP)])
        h.close()
    for v, out in zip(variants[1:], outs[1:]):
        for a, b in zip(outs[0], out):
            np.testing.assert_array_equal(a, b, err_msg=str(v))
    bt = outs[0][7]
    report(f"tier_splits[{name},T={T}]", batch=x0.shape[0], variants=len(variants), rollouts=int(bt.sum()),
           failed=int((outs[0][5] == abi.STATUS_LINESEARCH_FAILED).sum()))


# ------------------------------------------------------------------ lane-boundary constraints
def test_polyline2_signed_distance_constraint_on_the_device(product, oracle, oracle64):
    """ILQG_CONSTRAINT_POLYLINE2_SIGNED_DISTANCE (src/polyline2_signed_distance_constraint.cpp:58-145) on the
    device: ThreePlayerIntersection with the six lane-boundary constraints the reference example
    constructs and leaves commented out (src/three_player_intersection_example.cpp:214-251), cars
    started up to 4 m off their lane centres so that boundaries are violated, multipliers made
    non-zero by one ilqg_al_update sweep.  Records (K_lq), the LQ solve, the linesearch (merit
    kernel) and the multiplier update are compared with the oracle stage by stage."""
    desc, _ = problems.three_player_intersection(lane_constraints=True)
    B = 16
    x0 = problems.three_player_intersection_x0_batch(B, 77)
    rng = np.random.default_rng(5)
    x0[:, 0] += rng.uniform(-4.0, 4.0, B).astype(np.float32)    # P1 across its lane (x)
    x0[:, 6] += rng.uniform(-4.0, 4.0, B).astype(np.float32)    # P2
    x0[:, 13] += rng.uniform(-4.0, 4.0, B).astype(np.float32)   # P3 across its lane (y)
    hs = []
    for lib in (product, oracle, oracle64):
        h = abi.Handle(lib, desc, problems.three_player_intersection_params(), B, 0)
        h.upload_x0(x0)
        h.solve_begin()
        hs.append(h)
    c, o, o64 = hs
    assert c.layout.num_constraints == 12  # 6 proximity + 6 lane boundaries
    good = np.ones(B, bool)
    compared = []

    def check(what, tol=STAGE_TOL, label=""):
        nonlocal good
        a, b, b64 = c.download(what), o.download(what), o64.download(what)
        good = good & wellposed(b, b64)
        cond = None if "record" in label else row_distance(b, b64)
        compared.append(close(a, b, tol=tol, rows=good, what=f"{label} field {what}", cond=cond))

    for sweep in range(2):
        for h in hs:
            h.linearize_quadraticize()
        for what in (abi.QUAD_Q, abi.QUAD_L):
            check(what, label=f"sweep{sweep} record")
        for h in hs:
            h.lq_backward()
        for what in (abi.LQ_PS, abi.LQ_ALPHAS):
            check(what, tol=1e-3, label=f"sweep{sweep} LQ")
        for h in hs:
            h.linesearch()
        same = np.ones(B, bool)
        for what in (abi.STATUS, abi.ITERS, abi.BACKTRACKS):
            same &= c.download(what) == o.download(what)
        good = good & same & (o.download(abi.BACKTRACKS) == o64.download(abi.BACKTRACKS))
        for what in (abi.XS, abi.MERIT):
            check(what, tol=1e-3, label=f"sweep{sweep} linesearch")
        for h in hs:
            h.al_update()  # lambda <- max(0, lambda + mu g) at the current operating point
        lam_c, lam_o = c.download(abi.LAMBDAS), o.download(abi.LAMBDAS)
        compared.append(close(lam_c, lam_o, tol=1e-3, atol=1e-4, rows=good, what=f"sweep{sweep} lambdas"))
        # the lane multipliers are really in play: some boundary constraint is violated in most games
        lanes_active = (lam_o.reshape(B, 12, -1)[:, [0, 1, 4, 5, 8, 9]] > 0).any(axis=(1, 2))
        assert lanes_active.sum() >= B // 2, lanes_active
    report("polyline2_signed_distance_constraint", batch=B, wellposed_at_end=good.sum(), min_rows_compared=min(compared))
    assert good.sum() >= B // 2 and min(compared) >= B // 2
    for h in hs:
        h.close()
