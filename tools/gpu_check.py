"""Development battery: run the CUDA library and the oracle side by side and print the
per-stage differences.  Run on the GPU box (gpurun -- python tools/gpu_check.py)."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ilqgames_b200 import _abi as abi, problems  # noqa: E402

product = abi.product_library()
oracle = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so"))
oracle64 = abi.Library(os.path.join(REPO, "oracle", "_build", "libilqg_oracle64.so"))


def rowdiff(name, a, b):
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    err = np.abs(a - b).max(axis=1)
    print(f"  {name:18s} per-instance max|d| " + " ".join(f"{e:.2e}" for e in err) + "  | scale " +
          " ".join(f"{v:.1e}" for v in np.abs(b).max(axis=1)))


def diff(name, a, b, ref64=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    bad = ~np.isfinite(a) | ~np.isfinite(b)
    d = np.abs(a - b)
    d[bad & (np.isnan(a) == np.isnan(b))] = 0
    msg = f"  {name:18s} max|d|={np.nanmax(d):.3e} rel={np.nanmax(d) / scale:.3e} scale={scale:.3e}"
    if ref64 is not None:
        r = np.asarray(ref64, np.float64)
        msg += f" | cuda-f64={np.abs(a - r).max():.3e} oracle32-f64={np.abs(b - r).max():.3e}"
    print(msg)


def stage_compare(title, desc, params, x0, iters=3, fields=None):
    print(f"== {title}  B={x0.shape[0]}")
    hs = {}
    for name, lib in (("cuda", product), ("oracle", oracle), ("f64", oracle64)):
        h = abi.Handle(lib, desc, params, x0.shape[0], 0)
        h.upload_x0(x0)
        hs[name] = h
    for h in hs.values():
        h.solve_begin()
    for f, nm in ((abi.XS, "xs0"), (abi.US, "us0"), (abi.TOTAL_COSTS, "costs0")):
        diff(nm, hs["cuda"].download(f), hs["oracle"].download(f), hs["f64"].download(f))
    for it in range(iters):
        print(f" -- iteration {it + 1}")
        for h in hs.values():
            h.linearize_quadraticize()
        for f, nm in ((abi.LIN_A, "A"), (abi.LIN_B, "B"), (abi.QUAD_Q, "Q"), (abi.QUAD_L, "l"),
                      (abi.QUAD_R, "R"), (abi.QUAD_RGRAD, "r")):
            diff(nm, hs["cuda"].download(f), hs["oracle"].download(f), hs["f64"].download(f))
        for h in hs.values():
            h.lq_backward()
        for f, nm in ((abi.LQ_PS, "lqP"), (abi.LQ_ALPHAS, "lqAlpha"), (abi.DELTA_XS, "dxs"),
                      (abi.EXPECTED_DECREASE, "exp_decrease")):
            diff(nm, hs["cuda"].download(f), hs["oracle"].download(f), hs["f64"].download(f))
        for h in hs.values():
            h.linesearch()
        for f, nm in ((abi.XS, "xs"), (abi.US, "us"), (abi.ALPHAS, "alphas"), (abi.MERIT, "merit"),
                      (abi.STEP, "step"), (abi.TOTAL_COSTS, "costs")):
            diff(nm, hs["cuda"].download(f), hs["oracle"].download(f), hs["f64"].download(f))
        if x0.shape[0] <= 8:
            for f, nm in ((abi.LQ_PS, "lqP"), (abi.LQ_ALPHAS, "lqAlpha"), (abi.EXPECTED_DECREASE, "ed"),
                          (abi.XS, "xs"), (abi.MERIT, "merit")):
                rowdiff(nm, hs["cuda"].download(f), hs["oracle"].download(f))
        for f, nm in ((abi.STATUS, "status"), (abi.ITERS, "iters"), (abi.BACKTRACKS, "backtracks"),
                      (abi.TIME_OF_EXTREME, "t_extreme")):
            a, b = hs["cuda"].download(f), hs["oracle"].download(f)
            print(f"  {nm:18s} equal={np.array_equal(a, b)} cuda={a.ravel()[:8]} oracle={b.ravel()[:8]}")
    for h in hs.values():
        h.close()


def lq_test():
    print("== LQ test system (test/test_lq_solver.cpp)")
    sys.path.insert(0, os.path.join(REPO))
    from tests.test_oracle_pins import lq_test_system
    desc = problems.lq_only(100, 2, [1, 1], cross_pairs=[(0, 1), (1, 0)])
    for nominal in (0.0, 0.5):
        outs = {}
        for name, lib in (("cuda", product), ("oracle", oracle), ("f64", oracle64)):
            h = abi.Handle(lib, desc, abi.SolverParams.defaults(), 1)
            arrs, _ = lq_test_system(nominal)
            h.upload_lq(**arrs)
            h.lq_backward()
            outs[name] = (h.download(abi.LQ_PS), h.download(abi.LQ_ALPHAS))
            h.close()
        diff(f"P nominal={nominal}", outs["cuda"][0], outs["oracle"][0], outs["f64"][0])
        diff(f"alpha nominal={nominal}", outs["cuda"][1], outs["oracle"][1], outs["f64"][1])
        print("   P[0] cuda", outs["cuda"][0][0, 0], " alpha[0]", outs["cuda"][1][0, 0])


def full_solve(title, desc, params, x0, iters):
    print(f"== full iterate {title} B={x0.shape[0]} iters={iters}")
    outs = {}
    for name, lib in (("cuda", product), ("oracle", oracle)):
        h = abi.Handle(lib, desc, params, x0.shape[0], 0)
        h.upload_x0(x0)
        t = time.time()
        h.solve_begin()
        h.solve(chunk=iters)
        h.synchronize()
        dt = time.time() - t
        outs[name] = {f: h.download(f) for f in (abi.XS, abi.US, abi.STATUS, abi.ITERS, abi.BACKTRACKS,
                                                 abi.MERIT, abi.TOTAL_COSTS)}
        print(f"  {name}: {dt * 1e3:.1f} ms, launches={h.kernel_launches()}")
        h.close()
    c, o = outs["cuda"], outs["oracle"]
    same = (c[abi.STATUS] == o[abi.STATUS]) & (c[abi.ITERS] == o[abi.ITERS]) & (
        c[abi.BACKTRACKS] == o[abi.BACKTRACKS])
    print(f"  control flow identical for {same.sum()}/{len(same)} instances; "
          f"status hist cuda={np.bincount(c[abi.STATUS], minlength=6)} oracle={np.bincount(o[abi.STATUS], minlength=6)}")
    if same.any():
        diff("xs (same flow)", c[abi.XS][same], o[abi.XS][same])
        diff("us (same flow)", c[abi.US][same], o[abi.US][same])
        diff("merit (same flow)", c[abi.MERIT][same], o[abi.MERIT][same])
    if (~same).any():
        idx = np.nonzero(~same)[0][:5]
        print("  differing instances", idx, "backtracks cuda", c[abi.BACKTRACKS][idx], "oracle",
              o[abi.BACKTRACKS][idx], "merit", c[abi.MERIT][idx], o[abi.MERIT][idx])


if __name__ == "__main__":
    which = sys.argv[1:] or ["lq", "c1", "c3", "c4", "full"]
    if "lq" in which:
        lq_test()
    if "c1" in which:
        desc, _ = problems.three_player_intersection()
        stage_compare("ThreePlayerIntersection", desc, problems.three_player_intersection_params(),
                      problems.three_player_intersection_x0_batch(8, 1024))
    if "c3" in which:
        desc, _ = problems.roundabout_merging()
        stage_compare("RoundaboutMerging", desc, problems.roundabout_params(),
                      problems.roundabout_x0_batch(8, 4096), iters=2)
    if "c4" in which:
        desc, _ = problems.air_3d()
        stage_compare("Air3D", desc, problems.air_3d_params(), problems.air_3d_x0_grid(4)[:8], iters=2)
    if "golden" in which:
        for name, build, params in (("three_player_intersection", problems.three_player_intersection,
                                     problems.three_player_intersection_params),
                                    ("roundabout_merging", problems.roundabout_merging, problems.roundabout_params)):
            g = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
            desc, _ = build()
            stage_compare("golden " + name, desc, params(), g["x0"], iters=2)
    if "full" in which:
        desc, _ = problems.three_player_intersection()
        full_solve("ThreePlayerIntersection", desc,
                   problems.three_player_intersection_params(max_solver_iters=10),
                   problems.three_player_intersection_x0_batch(64, 1024), 10)
        desc, _ = problems.roundabout_merging()
        full_solve("RoundaboutMerging", desc, problems.roundabout_params(max_solver_iters=10),
                   problems.roundabout_x0_batch(32, 4096), 10)
        desc, _ = problems.air_3d()
        full_solve("Air3D", desc, problems.air_3d_params(max_solver_iters=10),
                   problems.air_3d_x0_grid(6), 10)
