"""The C++ host classes (include/ilqgames/**: Problem, ILQSolver, LQFeedbackSolver,
AugmentedLagrangianSolver, OperatingPoint, Strategy, ...) over the C ABI.

tests/cpp/host_api_test.cpp is written the way the reference's tests and executables use the
API; here it is compiled with g++, linked against an implementation of ilqg.h (the CPU oracle
for `-m "not gpu"`, libilqg_b200.so on the GPU box), run, and its dumped results are compared
with the same computations driven through ctypes.  Where /root/reference exists, the reference's
own src/three_player_intersection_example.cpp is compiled byte-for-byte unchanged against these
headers and must describe the same problem."""
import os
import subprocess

import numpy as np
import pytest

from ilqgames_b200 import _abi as abi, al, problems

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
SRC = os.path.join(REPO, "tests", "cpp", "host_api_test.cpp")


def _build(tmp_path, lib_path, dropin=False):
    exe = str(tmp_path / ("host_api_dropin" if dropin else "host_api_test"))
    libdir, libfile = os.path.split(lib_path)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(REPO, "include"), SRC, "-o", exe,
           "-L" + libdir, "-l" + libfile[3:-3], "-Wl,-rpath," + libdir]
    if dropin:
        # the reference's example sources and the three free functions they call, byte for byte;
        # compat/ supplies <glog/logging.h> and <gflags/gflags.h>, which the image does not have
        cmd[5:5] = ["-DDROPIN_REFERENCE_EXAMPLE", "-I" + os.path.join(REPO, "include", "ilqgames", "b200", "compat"),
                    "-I" + os.path.join(REFERENCE, "include")] + [
            os.path.join(REFERENCE, "src", f + ".cpp") for f in (
                "three_player_intersection_example", "roundabout_merging_example", "roundabout_lane_center",
                "initialize_along_route", "air_3d_example", "draw_shapes", "three_player_overtaking_example",
                "two_player_collision_example", "two_player_collision_avoidance_reachability_example",
                "three_player_collision_avoidance_reachability_example", "one_player_reachability_example",
                "dubins_origin_example", "two_player_reachability_example",
                "modified_air_3d_example", "modified_three_player_intersection_example", "skeleton_example",
                "three_player_intersection_reachability_example")]
    subprocess.run(cmd, check=True)
    return exe


def _run(exe, tmp_path):
    out = str(tmp_path / "out.bin")
    env = dict(os.environ, ILQGAMES_LOG_DIR=str(tmp_path / "logs"))
    res = subprocess.run([exe, out], capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all checks passed" in res.stdout
    arrays = {}
    with open(out, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            name, n = line.split()
            arrays[name.decode()] = np.frombuffer(f.read(4 * int(n)), dtype=np.float32)
    return arrays


def _params(**over):
    p = problems.three_player_intersection_params(**over)
    return p


def _flat_op(xs, us):
    return np.concatenate([xs, us], axis=-1).reshape(-1)


def _check_against_ctypes(lib, got, exact=True):
    desc, x0 = problems.three_player_intersection()
    # Problem -> descriptor: byte-identical to the Python builder
    assert np.array_equal(np.frombuffer(bytes(desc), dtype=np.uint32), got["desc_own"].view(np.uint32))
    assert np.array_equal(x0, got["x0_own"])
    # ... and with the lane boundaries on (Polyline2SignedDistanceConstraint::Describe)
    desc_l, _ = problems.three_player_intersection(lane_constraints=True)
    assert np.array_equal(np.frombuffer(bytes(desc_l), dtype=np.uint32), got["desc_lanes"].view(np.uint32))
    same = np.array_equal if exact else (lambda a, b: np.allclose(a, b, rtol=1e-3, atol=1e-3))

    # ILQSolver::Solve, max_solver_iters = 4
    h = abi.Handle(lib, desc, _params(max_solver_iters=4), 1, 0)
    h.upload_x0(x0[None])
    h.solve_begin()
    h.solve(chunk=1)
    assert int(got["ilq_num_iterates"][0]) == 1 + int(h.download(abi.ITERS)[0])
    assert same(_flat_op(h.download(abi.XS)[0], h.download(abi.US)[0]), got["ilq_final_op"])
    assert same(h.download(abi.TOTAL_COSTS)[0], got["ilq_total_costs"])
    Ps, alphas = h.download(abi.PS)[0], h.download(abi.ALPHAS)[0]       # [T][M][n], [T][M]
    T, M, n = Ps.shape
    strat = []
    for k in range(T):
        for i in range(3):
            strat += [Ps[k, 2 * i:2 * i + 2].reshape(-1), alphas[k, 2 * i:2 * i + 2]]
    assert same(np.concatenate(strat), got["ilq_final_strategies"])

    # SolveBatch games 1 and 2
    x0s = np.tile(x0, (3, 1))
    x0s[1, 0] += 1.0
    x0s[2, 7] -= 2.0
    hb = abi.Handle(lib, desc, _params(max_solver_iters=4), 3, 0)
    hb.upload_x0(x0s)
    hb.solve_begin()
    hb.solve()
    xs, us = hb.download(abi.XS), hb.download(abi.US)
    assert same(_flat_op(xs[1], us[1]), got["batch_op_1"])
    assert same(_flat_op(xs[2], us[2]), got["batch_op_2"])

    # AugmentedLagrangianSolver::Solve, NumIterates cap 30
    p = _params()
    p.max_solver_iters = p.unconstrained_solver_max_iters
    ha = abi.Handle(lib, desc, p, 1, 0)
    ha.upload_x0(x0[None])
    out = al.solve_augmented_lagrangian(ha, 30, p.constraint_error_tolerance)
    inner_solves, iterates, success = got["al_stats"]
    if exact:
        assert (int(inner_solves), int(iterates), int(success)) == (out.rounds, int(out.iterates[0]), int(out.success[0]))
    assert same(_flat_op(out.xs[0], out.us[0]), got["al_final_op"])
    # Problem::SetUpNextRecedingHorizon on the unconstrained variant, then the next solve
    desc_u, _ = problems.three_player_intersection(with_constraints=False)
    hr = abi.Handle(lib, desc_u, _params(max_solver_iters=3), 1, 0)
    hr.upload_x0(x0[None])
    hr.solve_begin()
    hr.solve(chunk=1)
    hr.overwrite_solution()
    xm = hr.download(abi.XS)[0, 3].copy()
    xm[0] += np.float32(0.05)
    xm[7] -= np.float32(0.03)
    # Dynamics()->Integrate(0.33, 0.81, x, plan, strategies) -> ilqg_integrate_plan
    assert same(hr.integrate_plan(xm[None], 0.33, 0.81)[0], got["ip_out"])
    new_t0 = hr.setup_next_receding_horizon(xm[None], 0.33, 0.25)
    assert np.float32(new_t0) == got["rh_t0"][0]
    assert same(hr.download(abi.X0)[0], got["rh_x0"])
    assert same(_flat_op(hr.download(abi.WARM_XS)[0], hr.download(abi.WARM_US)[0]), got["rh_op"])
    hr.solve_begin()
    hr.solve(chunk=1)
    assert same(_flat_op(hr.download(abi.XS)[0], hr.download(abi.US)[0]), got["rh_next_op"])
    # RecedingHorizonSimulator with a scripted clock: six solver calls, start times increasing
    sim = got["sim_summary"].reshape(6, 3 + 2 * 16)
    assert np.all(np.diff(sim[:, 0]) > 0) and sim[0, 0] == 0 and np.all(sim[:, 1] >= 1)
    # SolutionSplicer::Splice: the host class against the reference's own (tests/golden/ref_splice.npz)
    g = np.load(os.path.join(REPO, "tests", "golden", "ref_splice.npz"))
    for c in range(3):
        assert np.float32(g[f"splice{c}_t0"]) == got[f"splice{c}_t0"][0]
        for key in ("xs", "us", "alphas", "P00"):
            assert np.array_equal(g[f"splice{c}_{key}"].reshape(-1), got[f"splice{c}_{key}"]), (c, key)
    # SURVEY Appendix B known answer of the LQ test system
    np.testing.assert_allclose(got["lq_P1_k0"], [0.915024, 1.632025], atol=2e-5)
    np.testing.assert_allclose(got["lq_P2_k0"], [0.0145362, 0.0206234], atol=2e-5)


def _check_saved_log(tmp_path, got):
    """SolverLog::Save, the reference's on-disk format (src/solver_log.cpp:113-171)."""
    root = tmp_path / "logs" / "host_api_test"
    n_iter = int(got["ilq_num_iterates"][0])
    assert sorted(os.listdir(root), key=int) == [str(i) for i in range(n_iter)]
    last = root / str(n_iter - 1)
    assert sorted(os.listdir(last)) == ["costs.txt", "cumulative_runtimes.txt", "t0.txt", "u0.txt", "u1.txt",
                                        "u2.txt", "xs.txt"]
    op = got["ilq_final_op"].reshape(100, 22)
    xs = np.loadtxt(last / "xs.txt")
    us = np.concatenate([np.loadtxt(last / f"u{i}.txt") for i in range(3)], axis=1)
    np.testing.assert_allclose(xs, op[:, :16], rtol=6e-6, atol=1e-30)   # 6 significant digits
    np.testing.assert_allclose(us, op[:, 16:], rtol=6e-6, atol=1e-30)
    np.testing.assert_allclose(np.loadtxt(last / "costs.txt"), got["ilq_total_costs"], rtol=6e-6)
    assert float(open(last / "t0.txt").read()) == 0.0
    assert os.listdir(tmp_path / "logs" / "host_api_test_last") == [str(n_iter - 1)]
    # Eigen-style row: single-space separated, every cell padded to the widest
    first = open(last / "xs.txt").readline().rstrip("\n")
    cells = first.split()
    assert len(first) == len(cells) * max(map(len, cells)) + len(cells) - 1


def test_cpp_host_classes_on_the_oracle(oracle, tmp_path):
    got = _run(_build(tmp_path, oracle.path), tmp_path)
    _check_against_ctypes(oracle, got)
    _check_saved_log(tmp_path, got)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is not on this machine")
def test_reference_example_source_drops_in_unchanged(oracle, tmp_path):
    """north_star: 'an example like ThreePlayerIntersectionExample drops in unchanged'."""
    got = _run(_build(tmp_path, oracle.path, dropin=True), tmp_path)
    assert np.array_equal(got["desc_reference"].view(np.uint32), got["desc_own"].view(np.uint32))
    # RoundaboutMergingExample, Air3DExample, ThreePlayerOvertakingExample,
    # TwoPlayerCollisionExample: the descriptors DescribeProblem builds from the
    # reference's own example code equal the ones problems.py writes out by hand, bit for bit
    for tag, build in (("roundabout", problems.roundabout_merging), ("air3d", problems.air_3d),
                       ("overtaking", problems.three_player_overtaking),
                       ("collision", problems.two_player_collision),
                       ("reachability2", problems.two_player_collision_avoidance_reachability),
                       ("reachability3", problems.three_player_collision_avoidance_reachability),
                       ("reachability1", problems.one_player_reachability),
                       ("dubins_origin", problems.dubins_origin),
                       ("reachability_2p", problems.two_player_reachability),
                       ("modified_air3d", problems.modified_air_3d)):
        desc, x0 = build()
        mine = np.frombuffer(bytes(desc), dtype=np.uint32)
        theirs = got["desc_" + tag].view(np.uint32)
        assert np.array_equal(mine, theirs), (tag, np.nonzero(mine != theirs)[0][:8])
        assert np.array_equal(np.asarray(x0, np.float32), got["x0_" + tag]), tag
    assert np.array_equal(got["x0_reference"], got["x0_own"])
    # and no reference header other than the example's own declaration took part in the build
    deps = subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(REPO, "include"),
                           "-I" + os.path.join(REFERENCE, "include"), "-M",
                           os.path.join(REFERENCE, "src", "three_player_intersection_example.cpp")],
                          capture_output=True, text=True, check=True).stdout.split()
    from_ref = sorted(d for d in deps if d.startswith(REFERENCE))
    assert from_ref == [os.path.join(REFERENCE, "include/ilqgames/examples/three_player_intersection_example.h"),
                        os.path.join(REFERENCE, "src/three_player_intersection_example.cpp")]


@pytest.mark.gpu
def test_cpp_host_classes_on_the_gpu(product, tmp_path):
    got = _run(_build(tmp_path, product.path), tmp_path)
    _check_against_ctypes(product, got, exact=True)


def test_example_programs_build_and_run_on_the_oracle(oracle, tmp_path):
    """examples/cpp: the GUI-less counterparts of exec/three_player_intersection and
    exec/receding_horizon_example, linked against the CPU oracle here."""
    subprocess.run(["make", "-C", os.path.join(REPO, "examples", "cpp"), "ILQG_LIB=" + oracle.path,
                    "OUT=" + str(tmp_path)], check=True, capture_output=True)
    env = dict(os.environ, ILQGAMES_LOG_DIR=str(tmp_path / "logs"))
    res = subprocess.run([str(tmp_path / "three_player_intersection"), "4", "example_run"], capture_output=True,
                         text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "AugmentedLagrangianSolver:" in res.stdout and "ILQSolver::SolveBatch: 4 games" in res.stdout
    assert os.path.isdir(tmp_path / "logs" / "example_run")
    res = subprocess.run([str(tmp_path / "receding_horizon"), "1.0", "2.0"], capture_output=True, text=True,
                         env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    calls = [ln for ln in res.stdout.splitlines() if ln.strip().startswith("call ")]
    # a generous planner_runtime: the simulator CHECKs that no solve outlasts it, on a loaded host too
    assert len(calls) >= 2 and "horizon starts at t0 = 0.00 s" in calls[0]
