"""AugmentedLagrangianSolver::Solve (reference src/augmented_lagrangian_solver.cpp:72-210) for a
batch of games, on top of the C ABI (include/ilqg.h: ilqg_al_begin / ilqg_al_advance).

Every game runs its own outer loop on the device -- own multipliers, own mu, own iterate count,
own exit test -- while the host only sequences the rounds: one inner ILQSolver::Solve for all
games still in their loop, then one ilqg_al_advance.  Host traffic per round: two integers.

The reference's wall-clock gating is the caller's business: pass `max_rounds`."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi as abi

# include/ilqgames/solver/solver_params.h defaults
DEFAULT_MAX_SOLVER_ITERS = 1000
DEFAULT_CONSTRAINT_ERROR_TOLERANCE = 1e-1


@dataclass
class ALResult:
    rounds: int                 # inner solves issued (1 = unconstrained or satisfied at once)
    success: np.ndarray         # int32 [B]  AugmentedLagrangianSolver::Solve's *success
    iterates: np.ndarray        # int32 [B]  log->NumIterates()
    max_constraint_error: np.ndarray  # float32 [B] of the last multiplier sweep
    xs: np.ndarray              # final operating point / strategies (log->FinalOperatingPoint(),
    us: np.ndarray              # log->FinalStrategies()), [B][T][n], [B][T][M], [B][T][M][n], [B][T][M]
    Ps: np.ndarray
    alphas: np.ndarray


def solve_augmented_lagrangian(h: abi.Handle, max_solver_iters: int = DEFAULT_MAX_SOLVER_ITERS,
                               constraint_error_tolerance: float = DEFAULT_CONSTRAINT_ERROR_TOLERANCE,
                               reset_problem: bool = True, reset_lambdas: bool = True,
                               reset_mu: bool = True, initial_warmstart=None,
                               max_rounds: int = 10_000, chunk: int = 4) -> ALResult:
    """`h`'s ilqg_solver_params.max_solver_iters must be the INNER solver's cap
    (SolverParams::unconstrained_solver_max_iters); `max_solver_iters` here is the AL loop's
    NumIterates cap (augmented_lagrangian_solver.cpp:109).  The final iterate of each game is
    what ILQG_XS / ILQG_US / ILQG_PS / ILQG_ALPHAS download afterwards."""
    h.al_begin(max_solver_iters, constraint_error_tolerance)
    rounds = 0
    while rounds < max_rounds:
        h.solve_begin()
        h.solve(chunk=chunk)
        rounds += 1
        if h.al_advance() == 0:
            break
    out = ALResult(rounds, h.download(abi.AL_SUCCESS), h.download(abi.AL_ITERATES),
                   h.download(abi.MAX_CONSTRAINT_ERROR), h.download(abi.XS), h.download(abi.US),
                   h.download(abi.PS), h.download(abi.ALPHAS))
    # :192-207 -- the Problem gets its initial solution back and the multipliers their defaults,
    # unless the caller keeps them (SolverParams::reset_problem / reset_lambdas / reset_mu, independent
    # flags as in the reference).  ilqg_reset also ends the AL solve.
    mask = 0
    if reset_lambdas:
        mask |= h.RESET_LAMBDAS
    if reset_mu:
        mask |= h.RESET_MU
    if reset_problem and initial_warmstart is None:
        mask |= h.RESET_SOLUTION      # Problem::Initialize's zero operating point and strategies
    if mask:
        h.reset(mask)
    if reset_problem and initial_warmstart is not None:
        h.upload_warmstart(*initial_warmstart)
    return out
