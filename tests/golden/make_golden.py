"""Generate the golden iLQ trajectories in tests/golden/*.npz from the CPU oracle.

The reference holds no golden vectors for ILQSolver iterates (SURVEY.md sections 4, 8c) and
cannot be built here, so these fixtures come from the restatement in oracle/ilqg_oracle.cpp
(itself pinned by tests/test_oracle_pins.py).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi, problems  # noqa: E402

CASES = {
    # name: (descriptor builder, params builder, x0 batch, iterations)
    "three_player_intersection": (problems.three_player_intersection,
                                  problems.three_player_intersection_params,
                                  lambda: problems.three_player_intersection_x0_batch(12, 1024), 4),
    "roundabout_merging": (problems.roundabout_merging, problems.roundabout_params,
                           lambda: problems.roundabout_x0_batch(24, 4096), 2),
    "air_3d": (problems.air_3d, problems.air_3d_params, lambda: problems.air_3d_x0_grid(4)[:12], 3),
}


def run_case(lib, name, lib64=None):
    """Golden iterates from the fp32 oracle.  With lib64 (the fp64 build of the same code) also
    records `stable_<it>`: instances on which fp32 and fp64 still agree (same Armijo decisions,
    trajectories within 1e-3 relative).  Instances that are not stable sit on a knife edge of the
    reference algorithm itself (e.g. a car exactly on a polyline vertex, an Armijo test decided in
    the last bit, an ill-conditioned Riccati solve) and are not meaningful parity targets."""
    build, params, x0f, iters = CASES[name]
    desc, x0_example = build()
    x0 = x0f()
    x0[0] = x0_example  # instance 0 is the reference example's own initial state
    hs = []
    for l in ([lib] if lib64 is None else [lib, lib64]):
        h = abi.Handle(l, desc, params(max_solver_iters=iters), x0.shape[0])
        h.upload_x0(x0)
        h.solve_begin()
        hs.append(h)
    h = hs[0]
    out = {"x0": x0, "xs_0": h.download(abi.XS), "us_0": h.download(abi.US),
           "costs_0": h.download(abi.TOTAL_COSTS)}
    stable = np.ones(x0.shape[0], bool)
    for it in range(1, iters + 1):
        for hh in hs:
            hh.iterate(1)
        for key, what in (("xs", abi.XS), ("us", abi.US), ("merit", abi.MERIT), ("step", abi.STEP),
                          ("status", abi.STATUS), ("iters", abi.ITERS), ("backtracks", abi.BACKTRACKS),
                          ("costs", abi.TOTAL_COSTS), ("t_extreme", abi.TIME_OF_EXTREME)):
            out[f"{key}_{it}"] = h.download(what)
        if lib64 is not None:
            a, b = h.download(abi.XS).astype(np.float64), hs[1].download(abi.XS).astype(np.float64)
            with np.errstate(invalid="ignore"):
                rel = np.abs(a - b).max(axis=(1, 2)) / np.maximum(1.0, np.abs(b).max(axis=(1, 2)))
            stable &= np.isfinite(rel) & (rel < 1e-3)
            stable &= h.download(abi.BACKTRACKS) == hs[1].download(abi.BACKTRACKS)
            stable &= h.download(abi.STATUS) == hs[1].download(abi.STATUS)
            out[f"stable_{it}"] = stable.copy()
    for hh in hs:
        hh.close()
    return out


if __name__ == "__main__":
    import subprocess
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], check=True)
    lib = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so"))
    lib64 = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle64.so"))
    for name in CASES:
        out = run_case(lib, name, lib64)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print("wrote", name, {k: v.astype(int).tolist() for k, v in out.items() if k.startswith("stable")})
