"""A/B of the two record paths on one box (not part of the product):
  ILQG_RECORDS=dense    K_lq v3 -> dense records -> k_lq_backward_hw   (round 1)
  ILQG_RECORDS=compact  K_lq v4 -> compact records -> k_lq_backward_tc (round 2)
Stage outputs of one prologue + linearize/quadraticize + backward sweep are compared game by game,
then a few iLQ iterations are timed kernel by kernel.
usage: python tools/ab_records.py [c1|c3|c4|overtaking] [batch] [iters]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi, problems  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "c1"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3


def setup(batch, iters):
    if config == "c1":
        desc, _ = problems.three_player_intersection()
        params = problems.three_player_intersection_params(max_solver_iters=iters, disable_convergence_exit=1)
        x0 = problems.three_player_intersection_x0_batch(batch, 4096)
    elif config == "c3":
        desc, _ = problems.roundabout_merging()
        params = problems.roundabout_params(max_solver_iters=iters, disable_convergence_exit=1)
        x0 = problems.roundabout_x0_batch(batch, 4096)
    elif config == "overtaking":
        desc, _ = problems.three_player_overtaking()
        params = problems.three_player_overtaking_params(max_solver_iters=iters, disable_convergence_exit=1)
        x0 = problems.three_player_overtaking_x0_batch(batch, 18)
    else:
        desc, _ = problems.air_3d()
        params = problems.air_3d_params(max_solver_iters=iters, disable_convergence_exit=1)
        g = problems.air_3d_x0_grid(128)
        x0 = np.tile(g, (batch // len(g) + 1, 1))[:batch]
    return desc, params, x0


FIELDS = [("LIN_A", abi.LIN_A), ("LIN_B", abi.LIN_B), ("QUAD_Q", abi.QUAD_Q), ("QUAD_L", abi.QUAD_L),
          ("QUAD_R", abi.QUAD_R), ("QUAD_RGRAD", abi.QUAD_RGRAD), ("LQ_PS", abi.LQ_PS),
          ("LQ_ALPHAS", abi.LQ_ALPHAS), ("EXPECTED_DECREASE", abi.EXPECTED_DECREASE), ("DELTA_XS", abi.DELTA_XS)]


def stage(mode, warm_iters):
    os.environ["ILQG_RECORDS"] = mode
    desc, params, x0 = setup(batch, max(1, warm_iters))
    h = abi.Handle(abi.product_library(), desc, params, batch, 0)
    h.upload_x0(x0)
    h.solve_begin()
    if warm_iters:
        h.iterate(warm_iters)
    h.linearize_quadraticize()
    h.lq_backward()
    out = {name: h.download(what) for name, what in FIELDS}
    out["XS"] = h.download(abi.XS)
    h.close()
    return out


for warm in (0, 2):
    a, b = stage("dense", warm), stage("compact", warm)
    print(f"--- after {warm} iterations: compact path vs dense path ({batch} games)")
    # games whose trajectories already differ (linesearch decisions are chaotic) are not stage-comparable
    same = np.array([np.array_equal(a["XS"][g], b["XS"][g]) for g in range(batch)])
    print(f"  games with identical operating points: {same.sum()} / {batch}")
    for name, _ in FIELDS:
        x, y = a[name].reshape(batch, -1)[same], b[name].reshape(batch, -1)[same]
        scale = np.maximum(np.abs(x).max(axis=1), 1e-30)
        with np.errstate(invalid="ignore"):
            rel = np.abs(x - y).max(axis=1) / scale
        fin = np.isfinite(rel)
        exact = (x == y).all(axis=1).sum()
        if fin.any():
            print(f"  {name:18s} bit-identical games {exact:5d}; rel err median {np.median(rel[fin]):.2e} "
                  f"90% {np.quantile(rel[fin], 0.9):.2e} max {rel[fin].max():.2e}; non-finite {(~fin).sum()}")
        else:
            print(f"  {name:18s} no finite games")

# ---- timing, kernel by kernel ----
tb = int(os.environ.get("AB_TIMING_BATCH", "4096"))
for mode in ("dense", "compact"):
    os.environ["ILQG_RECORDS"] = mode
    batch_save = batch
    desc, params, _ = setup(1, 10)
    batch = tb
    desc, params, x0 = setup(tb, 10)
    batch = batch_save
    h = abi.Handle(abi.product_library(), desc, params, tb, 0)
    h.upload_x0(x0)
    for rep in range(2):
        h.reset(1)
        h.solve_begin()
        if rep == 1:
            h.profile(True)
        h.iterate(10)
        h.synchronize()
    prof = h.profile_read()
    h.profile(False)
    st = h.download(abi.STATUS)
    print(f"--- {mode}: batch {tb}, 10 iterations; status histogram {np.bincount(st, minlength=6).tolist()}, "
          f"backtracks {int(h.download(abi.BACKTRACKS).sum())}")
    for k, (ms, nl) in prof.items():
        if nl:
            print(f"  {k:24s} {ms / nl:8.4f} ms x {nl}")
    h.close()
