"""Precision study for a tensor-core K_bwd -- CPU only.

The backward sweep of LQFeedbackSolver (src/lq_feedback_solver.cpp:71-244) is re-run in numpy on
the records the oracle produces for the benchmark batch, with the dense products (B'Z, (B'Z)B,
(B'Z)A, F'ZF, Z beta, F'(...)) computed four ways:

  f64      double throughout (the yardstick)
  f32      float32 products (what the CUDA-core kernel does)
  tf32     operands rounded to TF32 (10-bit mantissa), fp32 accumulation: one mma per product
  tf32x3   operands split hi + lo, a.b ~ hi.hi + (hi.lo + lo.hi): three mma per product

The 6 x 6 solve, Gershgorin and the small P'R terms stay in float32 in every variant.  Reported:
norm-wise error of P and alpha against f64 per game (max over the horizon), the quantity the
parity tests bound by 1e-4 (tests/test_gpu_parity.py STAGE_TOL).

    python tools/tf32_study.py [games] [iterations before the records are taken]
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi, problems  # noqa: E402


def to_tf32(a):
    """Round float32 to TF32 (keep 10 mantissa bits), round-to-nearest-even."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def make_mm(mode):
    if mode == "f64":
        return lambda a, b: np.matmul(a.astype(np.float64), b.astype(np.float64))
    if mode == "f32":
        return lambda a, b: np.matmul(a.astype(np.float32), b.astype(np.float32))
    if mode == "tf32":
        return lambda a, b: np.matmul(to_tf32(a), to_tf32(b))
    if mode == "tf32x3":
        def mm(a, b):
            a, b = a.astype(np.float32), b.astype(np.float32)
            ah, bh = to_tf32(a), to_tf32(b)
            al, bl = to_tf32(a - ah), to_tf32(b - bh)
            return np.matmul(ah, bh) + (np.matmul(ah, bl) + np.matmul(al, bh))
        return mm
    raise ValueError(mode)


def backward(A, Bm, Q, l, R, r, udim, mode):
    """A [G][T][n][n], Bm [G][T][n][M], Q [G][T][N][n][n], l [G][T][N][n], R [G][T][N][m][m],
    r [G][T][N][m] (own-control pairs only) -> P [G][T][M][n], alpha [G][T][M]."""
    mm = make_mm(mode)
    dt = np.float64 if mode == "f64" else np.float32
    G, T, n, _ = A.shape
    N, M = len(udim), sum(udim)
    off = np.concatenate([[0], np.cumsum(udim)])
    A, Bm, Q, l, R, r = (x.astype(dt) for x in (A, Bm, Q, l, R, r))
    P = np.zeros((G, T, M, n), dt)
    al = np.zeros((G, T, M), dt)
    Z = Q[:, T - 1].copy()          # [G][N][n][n]
    zeta = l[:, T - 1].copy()       # [G][N][n]
    for k in range(T - 2, -1, -1):
        S = np.zeros((G, M, M), dt)
        Y = np.zeros((G, M, n + 1), dt)
        for i in range(N):
            Bi = Bm[:, k, :, off[i]:off[i + 1]]                       # [G][n][m]
            BiZi = mm(np.swapaxes(Bi, 1, 2), Z[:, i]).astype(dt)      # [G][m][n]
            S[:, off[i]:off[i + 1], :] = mm(BiZi, Bm[:, k]).astype(dt)
            S[:, off[i]:off[i + 1], off[i]:off[i + 1]] += R[:, k, i]
            Y[:, off[i]:off[i + 1], :n] = mm(BiZi, A[:, k]).astype(dt)
            Y[:, off[i]:off[i + 1], n] = mm(np.swapaxes(Bi, 1, 2), zeta[:, i, :, None]).astype(dt)[..., 0] + r[:, k, i]
        for c in range(M):          # Gershgorin, column-wise (:163-176)
            radius = np.abs(S[:, :, c]).sum(axis=1) - np.abs(S[:, c, c])
            low = S[:, c, c] - radius < dt(1e-3)
            S[low, c, c] += radius[low] + dt(1e-3)
        X = np.linalg.solve(S, Y).astype(dt)
        P[:, k], al[:, k] = X[:, :, :n], X[:, :, n]
        F = A[:, k] - sum(np.matmul(Bm[:, k, :, off[i]:off[i + 1]], P[:, k, off[i]:off[i + 1]]) for i in range(N))
        beta = -sum(np.matmul(Bm[:, k, :, off[i]:off[i + 1]], al[:, k, off[i]:off[i + 1], None])[..., 0] for i in range(N))
        F = F.astype(dt)
        beta = beta.astype(dt)
        Ft = np.swapaxes(F, 1, 2)
        for i in range(N):
            Pi, ai = P[:, k, off[i]:off[i + 1]], al[:, k, off[i]:off[i + 1]]
            Zb = mm(Z[:, i], beta[:, :, None]).astype(dt)[..., 0]
            zeta_i = mm(Ft, (zeta[:, i] + Zb)[:, :, None]).astype(dt)[..., 0] + l[:, k, i]
            Z_i = mm(Ft, mm(Z[:, i], F).astype(dt)).astype(dt) + Q[:, k, i]
            Ra = np.matmul(R[:, k, i], ai[:, :, None])[..., 0] - r[:, k, i]
            zeta[:, i] = zeta_i + np.matmul(np.swapaxes(Pi, 1, 2), Ra[:, :, None])[..., 0]
            Z[:, i] = Z_i + np.matmul(np.swapaxes(Pi, 1, 2), np.matmul(R[:, k, i], Pi))
    return P, al


def main():
    games = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lib = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so"))
    desc, _ = problems.three_player_intersection()
    h = abi.Handle(lib, desc, problems.three_player_intersection_params(max_solver_iters=max(warm, 1)), games)
    h.upload_x0(problems.three_player_intersection_x0_batch(games, 4096))
    h.solve_begin()
    if warm:
        h.solve(chunk=warm)
    h.linearize_quadraticize()
    A, Bm = h.download(abi.LIN_A), h.download(abi.LIN_B)
    Q, l = h.download(abi.QUAD_Q), h.download(abi.QUAD_L)
    T = A.shape[1]
    R = h.download(abi.QUAD_R).reshape(games, T, 3, 2, 2)
    r = h.download(abi.QUAD_RGRAD).reshape(games, T, 3, 2)
    h.lq_backward()
    P_or, a_or = h.download(abi.LQ_PS), h.download(abi.LQ_ALPHAS)
    P64, a64 = backward(A, Bm, Q, l, R, r, [2, 2, 2], "f64")

    def err(x, x64):
        x, x64 = x.reshape(games, -1).astype(np.float64), x64.reshape(games, -1)
        return np.abs(x - x64).max(axis=1) / np.maximum(np.abs(x64).max(axis=1), 1e-30)

    tame = np.isfinite(P64).reshape(games, -1).all(axis=1) & (np.abs(P64).reshape(games, -1).max(axis=1) < 1e6)
    print(f"records after a solve capped at {warm} iterations, {games} games ({int(tame.sum())} tame in f64), T = {T}")
    print(f"{'products':10s} {'finite':>6s} {'P median':>10s} {'P 90%':>10s} {'P max':>10s} {'a median':>10s} {'a 90%':>10s} {'a max':>10s}")
    with np.errstate(all="ignore"):
        rows = [("oracle", P_or, a_or)] + [(m,) + backward(A, Bm, Q, l, R, r, [2, 2, 2], m) for m in ("f32", "tf32x3", "tf32")]
    for name, P, a in rows:
        ok = tame & np.isfinite(P).reshape(games, -1).all(axis=1)
        eP, ea = err(P, P64)[ok], err(a, a64)[ok]
        q = lambda e: (np.median(e), np.quantile(e, 0.9), e.max())
        print(f"{name:10s} {int(ok.sum()):6d} " + " ".join(f"{v:10.2e}" for v in q(eP) + q(ea)))

if __name__ == "__main__":
    main()
