"""AugmentedLagrangianSolver::Solve as a per-game state machine (ilqg_al_begin / ilqg_al_advance,
reference src/augmented_lagrangian_solver.cpp:72-210) -- CPU tests on the oracle.

No reference test pins the AL loop, so the state machine is checked against
(1) the loop written out with the primitive calls in the shape of the reference's code, and
(2) the property that makes batching legitimate: a game's result does not depend on which other
games share its batch."""
import numpy as np
import pytest

from ilqgames_b200 import _abi as abi, al, problems


def _handle(lib, batch, x0, config="c1", **over):
    if config == "c1":
        desc, _ = problems.three_player_intersection()
        params = problems.three_player_intersection_params(**over)
    else:
        desc, _ = problems.roundabout_merging()
        params = problems.roundabout_params(**over)
    # the inner ILQSolver is built with max_solver_iters = unconstrained_solver_max_iters
    # (include/ilqgames/solver/augmented_lagrangian_solver.h:82-83)
    params.max_solver_iters = params.unconstrained_solver_max_iters
    h = abi.Handle(lib, desc, params, batch, 0)
    h.upload_x0(x0)
    return h


def _reference_shaped_loop(h, max_iterates, tol):
    """augmented_lagrangian_solver.cpp:72-190 for batch 1, written with the primitive ABI calls."""
    success = True
    h.solve_begin(); h.solve()
    ok = int(h.download(abi.STATUS)[0]) != abi.STATUS_LINESEARCH_FAILED
    iterates = 1 + int(h.download(abi.ITERS)[0]) - (0 if ok else 1)
    success &= ok
    max_err = np.inf
    rounds = 1
    while iterates < max_iterates and max_err > tol:
        h.al_update()
        max_err = float(h.download(abi.MAX_CONSTRAINT_ERROR)[0])
        if ok:
            h.overwrite_solution(False)
        h.solve_begin(); h.solve()
        rounds += 1
        ok = int(h.download(abi.STATUS)[0]) != abi.STATUS_LINESEARCH_FAILED
        if not ok:
            h.al_post_solve()
        success &= ok
        iterates += 1 + int(h.download(abi.ITERS)[0]) - (0 if ok else 1)
    if max_err > tol:
        success = False
    return dict(rounds=rounds, success=int(success), iterates=iterates, max_err=max_err,
                xs=h.download(abi.XS), lambdas=h.download(abi.LAMBDAS), mu=h.download(abi.MU))


@pytest.mark.parametrize("seed", [1024, 7])
def test_state_machine_equals_reference_shaped_loop(oracle, seed):
    x0 = problems.three_player_intersection_x0_batch(1, seed)
    ref = _reference_shaped_loop(_handle(oracle, 1, x0), 40, 0.1)
    h = _handle(oracle, 1, x0)
    out = al.solve_augmented_lagrangian(h, 40, 0.1, reset_problem=False, reset_lambdas=False, reset_mu=False)
    assert out.rounds == ref["rounds"]
    assert int(out.success[0]) == ref["success"]
    assert int(out.iterates[0]) == ref["iterates"]
    assert float(out.max_constraint_error[0]) == ref["max_err"]
    np.testing.assert_array_equal(out.xs, ref["xs"])
    np.testing.assert_array_equal(h.download(abi.LAMBDAS), ref["lambdas"])
    np.testing.assert_array_equal(h.download(abi.MU), ref["mu"])
    assert ref["rounds"] >= 2, "the constrained example must go around the outer loop"
    assert (h.download(abi.AL_STATE) == 2).all()


def test_games_are_independent_of_their_batch(oracle):
    x0 = problems.three_player_intersection_x0_batch(3, 99)
    together = al.solve_augmented_lagrangian(_handle(oracle, 3, x0), 30, 0.1)
    for b in range(3):
        alone = al.solve_augmented_lagrangian(_handle(oracle, 1, x0[b:b + 1]), 30, 0.1)
        assert int(alone.iterates[0]) == int(together.iterates[b])
        assert int(alone.success[0]) == int(together.success[b])
        np.testing.assert_array_equal(alone.xs[0], together.xs[b])
        np.testing.assert_array_equal(alone.alphas[0], together.alphas[b])
    # the games do leave the loop at different times, which is what the masks are for
    assert together.rounds >= 2


def test_unconstrained_problem_is_one_inner_solve(oracle):
    x0 = problems.roundabout_x0_batch(2, 5)
    h = _handle(oracle, 2, x0, config="c3")
    out = al.solve_augmented_lagrangian(h, 100, 0.1)
    assert out.rounds == 1
    status = h.download(abi.STATUS)
    np.testing.assert_array_equal(out.success, (status != abi.STATUS_LINESEARCH_FAILED).astype(np.int32))
    assert (out.iterates >= 1).all() and (out.iterates <= 11).all()


def test_finished_games_are_frozen_and_reset_releases_them(oracle):
    x0 = problems.three_player_intersection_x0_batch(2, 3)
    h = _handle(oracle, 2, x0)
    h.al_begin(1, 0.1)             # NumIterates cap of 1: every game stops after its first solve
    h.solve_begin(); h.solve()
    assert h.al_advance() == 0
    xs = h.download(abi.XS)
    status = h.download(abi.STATUS)
    h.solve_begin()                # must not touch finished games
    np.testing.assert_array_equal(h.download(abi.STATUS), status)
    np.testing.assert_array_equal(h.download(abi.XS), xs)
    h.reset(h.RESET_SOLVER)
    h.solve_begin()
    assert (h.download(abi.STATUS) == abi.STATUS_RUNNING).all()
    assert (h.download(abi.AL_STATE) == 0).all()


def test_multiplier_reset_defaults(oracle):
    x0 = problems.three_player_intersection_x0_batch(1, 11)
    h = _handle(oracle, 1, x0)
    out = al.solve_augmented_lagrangian(h, 25, 0.1)   # reset_* default to the reference's true
    assert out.rounds >= 2
    assert (h.download(abi.LAMBDAS) == 0).all() and (h.download(abi.MU) == 10.0).all()
    assert (h.download(abi.XS) == 0).all(), "reset_problem restores Problem::Initialize's zeros"
    assert np.abs(out.xs).max() > 0
