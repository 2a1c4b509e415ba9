"""Per-game error table of one backward sweep (not part of the product): CUDA vs fp32 oracle vs fp64
oracle on the stage-parity test's inputs, to see how the CUDA path's distance to the fp64 answer
compares with the fp32 oracle's own.  usage: python tools/stage_errors.py [config] [batch]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi, problems  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "three_player_intersection"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = {
    "three_player_intersection": (problems.three_player_intersection, problems.three_player_intersection_params,
                                  lambda b: problems.three_player_intersection_x0_batch(b, 1024)),
    "roundabout_merging": (problems.roundabout_merging, problems.roundabout_params,
                           lambda b: problems.roundabout_x0_batch(b, 4096)),
    "air_3d": (problems.air_3d, problems.air_3d_params, lambda b: problems.air_3d_x0_grid(8)[:b]),
}[name]
desc, _ = cfg[0]()
x0 = cfg[2](batch)
libs = {"cuda": abi.product_library(),
        "o32": abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so")),
        "o64": abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle64.so"))}
out = {}
for k, lib in libs.items():
    h = abi.Handle(lib, desc, cfg[1](), x0.shape[0], 0)
    h.upload_x0(x0)
    h.solve_begin()
    h.linearize_quadraticize()
    h.lq_backward()
    out[k] = {f: h.download(w).reshape(x0.shape[0], -1).astype(np.float64)
              for f, w in (("P", abi.LQ_PS), ("alpha", abi.LQ_ALPHAS), ("ed", abi.EXPECTED_DECREASE))}
    h.close()
for f in ("P", "alpha", "ed"):
    print(f"--- {name} {f}: per game  |cuda-o32|  |o32-o64|  |cuda-o64|   (relative to max|o64|)")
    s = np.maximum(np.abs(out["o64"][f]).max(axis=1), 1e-30)
    e1 = np.abs(out["cuda"][f] - out["o32"][f]).max(axis=1) / s
    e2 = np.abs(out["o32"][f] - out["o64"][f]).max(axis=1) / s
    e3 = np.abs(out["cuda"][f] - out["o64"][f]).max(axis=1) / s
    for g in range(x0.shape[0]):
        print(f"  game {g:3d}  {e1[g]:.2e}  {e2[g]:.2e}  {e3[g]:.2e}   ratio cuda/o32 error {e3[g] / max(e2[g], 1e-30):6.2f}")
    fin = np.isfinite(e3) & np.isfinite(e2)
    print(f"  median ratio {np.median((e3 / np.maximum(e2, 1e-30))[fin]):.2f}")
