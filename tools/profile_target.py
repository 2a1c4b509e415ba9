"""Short, torch-free target for ncu: one Solve() prologue + a few iLQ iterations at the bench
batch size.  usage: python tools/profile_target.py [batch] [iters] [c1|c3|overtaking|c4] [finish]
With a fifth argument the solve also runs to the end and a state is carried along the plan
(ilqg_integrate_plan), so compute-sanitizer sees those kernels too."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi, problems  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
config = sys.argv[3] if len(sys.argv) > 3 else "c1"
if config == "c1":
    desc, _ = problems.three_player_intersection()
    params = problems.three_player_intersection_params(max_solver_iters=iters, disable_convergence_exit=1)
    x0 = problems.three_player_intersection_x0_batch(batch, 4096)
elif config == "c3":
    desc, _ = problems.roundabout_merging()
    params = problems.roundabout_params(max_solver_iters=iters, disable_convergence_exit=1)
    x0 = problems.roundabout_x0_batch(batch, 4096)
elif config == "overtaking":   # n = 18: the warp-per-game K_bwd instance
    desc, _ = problems.three_player_overtaking()
    params = problems.three_player_overtaking_params(max_solver_iters=iters, disable_convergence_exit=1)
    x0 = problems.three_player_overtaking_x0_batch(batch, 18)
else:
    desc, _ = problems.air_3d()
    params = problems.air_3d_params(max_solver_iters=iters, disable_convergence_exit=1)
    import numpy as np
    g = problems.air_3d_x0_grid(128)
    x0 = np.tile(g, (batch // len(g) + 1, 1))[:batch]
h = abi.Handle(abi.product_library(), desc, params, batch, 0)
h.upload_x0(x0)
h.solve_begin()
h.iterate(iters)
h.synchronize()
if len(sys.argv) > 4:
    h.solve()
    h.overwrite_solution()
    h.integrate_plan(x0, 0.13, 0.52)
print("done", h.download(abi.ITERS).sum(), h.download(abi.BACKTRACKS).sum())
