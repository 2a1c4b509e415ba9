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
                                  lambda: problems.three_player_intersection_x0_batch(6, 1024), 4),
    "roundabout_merging": (problems.roundabout_merging, problems.roundabout_params,
                           lambda: problems.roundabout_x0_batch(4, 4096), 3),
    "air_3d": (problems.air_3d, problems.air_3d_params, lambda: problems.air_3d_x0_grid(3)[:6], 3),
}


def run_case(lib, name):
    build, params, x0f, iters = CASES[name]
    desc, x0_example = build()
    x0 = x0f()
    x0[0] = x0_example  # instance 0 is the reference example's own initial state
    h = abi.Handle(lib, desc, params(max_solver_iters=iters), x0.shape[0])
    h.upload_x0(x0)
    h.solve_begin()
    out = {"x0": x0, "xs_0": h.download(abi.XS), "us_0": h.download(abi.US),
           "costs_0": h.download(abi.TOTAL_COSTS)}
    for it in range(1, iters + 1):
        h.iterate(1)
        for key, what in (("xs", abi.XS), ("us", abi.US), ("merit", abi.MERIT), ("step", abi.STEP),
                          ("status", abi.STATUS), ("iters", abi.ITERS), ("backtracks", abi.BACKTRACKS),
                          ("costs", abi.TOTAL_COSTS), ("t_extreme", abi.TIME_OF_EXTREME)):
            out[f"{key}_{it}"] = h.download(what)
    h.close()
    return out


if __name__ == "__main__":
    import subprocess
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], check=True)
    lib = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so"))
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **run_case(lib, name))
        print("wrote", name)
