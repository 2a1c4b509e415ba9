"""Diagnostics for the widened examples (not part of the product): per game, after k iterations, the
distance of the CUDA path / fp32 oracle / fp64 oracle trajectories to each other, and the control flow."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi
import tests.test_gpu_parity as parity

names = sys.argv[1:] or ["modified_air_3d", "two_player_collision"]
libs = {"cuda": abi.product_library(),
        "o32": abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so")),
        "o64": abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle64.so"))}
for name in names:
    build, params, x0f = parity.CONFIGS[name]
    desc, _ = build(num_time_steps=100)
    x0 = x0f(8)
    for iters in (1, 2):
        out = {}
        for k, lib in libs.items():
            h = abi.Handle(lib, desc, params(max_solver_iters=iters), x0.shape[0], 0)
            h.upload_x0(x0)
            h.solve_begin()
            h.solve(chunk=iters)
            out[k] = {w: h.download(w) for w in (abi.XS, abi.US, abi.STATUS, abi.ITERS, abi.BACKTRACKS, abi.MERIT, abi.STEP)}
            h.close()
        print(f"--- {name}: after {iters} iteration(s)")
        for g in range(x0.shape[0]):
            s = max(np.abs(out["o64"][abi.XS][g]).max(), 1e-30)
            e = lambda a, b: np.abs(out[a][abi.XS][g].astype(np.float64) - out[b][abi.XS][g]).max() / s
            print(f"  game {g}: backtracks c/o32/o64 {out['cuda'][abi.BACKTRACKS][g]}/{out['o32'][abi.BACKTRACKS][g]}/{out['o64'][abi.BACKTRACKS][g]}"
                  f"  status {out['cuda'][abi.STATUS][g]}/{out['o32'][abi.STATUS][g]}/{out['o64'][abi.STATUS][g]}"
                  f"  xs err cuda-o32 {e('cuda','o32'):.2e} o32-o64 {e('o32','o64'):.2e} cuda-o64 {e('cuda','o64'):.2e}"
                  f"  merit {out['cuda'][abi.MERIT][g]:.6g}/{out['o32'][abi.MERIT][g]:.6g}/{out['o64'][abi.MERIT][g]:.6g}")
