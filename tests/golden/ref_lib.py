"""ctypes wrapper of oracle/_ref/libilqg_ref.so -- the reference's OWN sources (compiled by
`make -C oracle ref` against the Eigen/glog/gflags stand-ins in oracle/ref_shim) behind the few
entry points of oracle/ref_driver.cpp.  TEST INFRASTRUCTURE: used by make_ref_golden.py (fixture
generation, needs /root/reference) and by tests/test_ref_pins.py when the library is present."""
import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libilqg_ref.so")

INTERSECTION, ROUNDABOUT, AIR3D, OVERTAKING, COLLISION, REACHABILITY2, REACHABILITY3, REACHABILITY1, DUBINS_ORIGIN, REACHABILITY_2P, MODIFIED_AIR3D = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
MODIFIED_INTERSECTION, SKELETON, INTERSECTION_REACHABILITY = 11, 12, 13
ILQ, AL = 0, 1


class RefParams(C.Structure):
    _fields_ = [("convergence_tolerance", C.c_float), ("max_solver_iters", C.c_int32),
                ("linesearch", C.c_int32), ("initial_alpha_scaling", C.c_float),
                ("geometric_alpha_scaling", C.c_float), ("max_backtracking_steps", C.c_int32),
                ("expected_decrease_fraction", C.c_float), ("open_loop", C.c_int32),
                ("unconstrained_solver_max_iters", C.c_int32), ("geometric_mu_scaling", C.c_float),
                ("geometric_mu_downscaling", C.c_float),
                ("geometric_lambda_downscaling", C.c_float),
                ("constraint_error_tolerance", C.c_float)]

    @classmethod
    def from_abi(cls, p) -> "RefParams":
        """From an ilqgames_b200._abi.SolverParams (same field names)."""
        return cls(**{name: getattr(p, name) for name, _ in cls._fields_})


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefLibrary:
    def __init__(self, path: str = REF_LIB):
        self.lib = C.CDLL(path)

    def dims(self, which: int):
        v = [C.c_int(0) for _ in range(5)]
        assert self.lib.ilqg_ref_dims(which, *[C.byref(x) for x in v]) == 0
        return tuple(x.value for x in v)  # n, M, N, T, num_constraints

    def x0(self, which: int) -> np.ndarray:
        n = self.dims(which)[0]
        out = np.zeros(n, np.float32)
        assert self.lib.ilqg_ref_x0(which, _ptr(out)) == 0
        return out

    def solve(self, which: int, solver: int, x0, params: RefParams, mu0: float = 10.0,
              max_log: int = 64):
        n, M, N, T, nc = self.dims(which)
        x0 = np.ascontiguousarray(x0, np.float32)
        xs = np.zeros((max_log, T, n), np.float32)
        us = np.zeros((max_log, T, M), np.float32)
        Ps = np.zeros((T, M, n), np.float32)
        alphas = np.zeros((T, M), np.float32)
        costs = np.zeros(N, np.float32)
        lambdas = np.zeros((max(nc, 1), T), np.float32)
        mu = C.c_float(0)
        iterates, success = C.c_int(0), C.c_int(0)
        rc = self.lib.ilqg_ref_solve(which, solver, _ptr(x0), C.byref(params), C.c_float(mu0),
                                     max_log, _ptr(xs), _ptr(us), _ptr(Ps), _ptr(alphas),
                                     _ptr(costs), C.byref(iterates), C.byref(success),
                                     _ptr(lambdas), C.byref(mu))
        assert rc == 0
        k = min(iterates.value, max_log)
        return dict(xs=xs[:k], us=us[:k], Ps=Ps, alphas=alphas, costs=costs,
                    iterates=iterates.value, success=success.value, lambdas=lambdas[:nc],
                    mu=mu.value)

    def lin_quad(self, which: int, x0, params: RefParams, mu0: float = 10.0):
        n, M, N, T, _ = self.dims(which)
        x0 = np.ascontiguousarray(x0, np.float32)
        A = np.zeros((T, n, n), np.float32)
        B = np.zeros((T, n, M), np.float32)
        Q = np.zeros((T, N, n, n), np.float32)
        l = np.zeros((T, N, n), np.float32)
        m2 = self._sum_udim_sq(which)
        R = np.zeros((T, m2), np.float32)
        r = np.zeros((T, M), np.float32)
        rc = self.lib.ilqg_ref_lin_quad(which, _ptr(x0), C.byref(params), C.c_float(mu0), _ptr(A),
                                        _ptr(B), _ptr(Q), _ptr(l), _ptr(R), _ptr(r))
        assert rc in (0, 1)
        return dict(A=A, B=B, Q=Q, l=l, R=R, r=r, ok=rc == 0)

    def receding_horizon(self, which: int, x0, params: RefParams, x_meas, t: float,
                         planner_runtime: float):
        n, M, N, T, _ = self.dims(which)
        x0 = np.ascontiguousarray(x0, np.float32)
        x_meas = np.ascontiguousarray(x_meas, np.float32)
        x0_out = np.zeros(n, np.float32)
        xs, us = np.zeros((T, n), np.float32), np.zeros((T, M), np.float32)
        Ps, alphas = np.zeros((T, M, n), np.float32), np.zeros((T, M), np.float32)
        t0 = C.c_double(0)
        rc = self.lib.ilqg_ref_receding_horizon(which, _ptr(x0), C.byref(params), _ptr(x_meas),
                                                C.c_double(t), C.c_double(planner_runtime),
                                                _ptr(x0_out), _ptr(xs), _ptr(us), _ptr(Ps),
                                                _ptr(alphas), C.byref(t0))
        assert rc in (0, 1)
        return dict(x0=x0_out, xs=xs, us=us, Ps=Ps, alphas=alphas, t0=t0.value, ok=rc == 0)

    def integrate(self, which: int, plan: dict, x0, t0: float, t: float) -> np.ndarray:
        """MultiPlayerIntegrableSystem::Integrate(t0, t, x0, plan) of problem `which`'s dynamics;
        plan = dict(xs [S][n], us [S][M], Ps [S][M][n], alphas [S][M], t0)."""
        n = self.dims(which)[0]
        arrs = [np.ascontiguousarray(plan[k], np.float32) for k in ("xs", "us", "Ps", "alphas")]
        x0 = np.ascontiguousarray(x0, np.float32)
        out = np.zeros(n, np.float32)
        rc = self.lib.ilqg_ref_integrate(which, int(arrs[0].shape[0]), *[_ptr(a) for a in arrs],
                                         C.c_double(plan["t0"]), _ptr(x0), C.c_double(t0), C.c_double(t),
                                         _ptr(out))
        assert rc == 0
        return out

    def receding_from_plan(self, which: int, plan: dict, x_meas, t: float, planner_runtime: float):
        """Problem::OverwriteSolution(plan) + SetUpNextRecedingHorizon; the plan may be longer than
        the horizon (a spliced plan)."""
        n, M, N, T, _ = self.dims(which)
        arrs = [np.ascontiguousarray(plan[k], np.float32) for k in ("xs", "us", "Ps", "alphas")]
        x_meas = np.ascontiguousarray(x_meas, np.float32)
        x0_out = np.zeros(n, np.float32)
        xs, us = np.zeros((T, n), np.float32), np.zeros((T, M), np.float32)
        Ps, alphas = np.zeros((T, M, n), np.float32), np.zeros((T, M), np.float32)
        t0 = C.c_double(0)
        rc = self.lib.ilqg_ref_receding_from_plan(
            which, int(arrs[0].shape[0]), *[_ptr(a) for a in arrs], C.c_double(plan["t0"]), _ptr(x_meas),
            C.c_double(t), C.c_double(planner_runtime), _ptr(x0_out), _ptr(xs), _ptr(us), _ptr(Ps),
            _ptr(alphas), C.byref(t0))
        assert rc == 0, rc
        return dict(x0=x0_out, xs=xs, us=us, Ps=Ps, alphas=alphas, t0=t0.value)

    def splice(self, new_t0: float, max_steps: int = 128):
        """SolutionSplicer::Splice on the synthetic logs of oracle/ref_driver.cpp."""
        xs, us, al = (np.zeros((max_steps, 3), np.float32) for _ in range(3))
        p00 = np.zeros((max_steps, 2), np.float32)
        t0 = C.c_double(0)
        k = self.lib.ilqg_ref_splice(C.c_double(new_t0), max_steps, _ptr(xs), _ptr(us), _ptr(al), _ptr(p00),
                                     C.byref(t0))
        return dict(xs=xs[:k], us=us[:k], alphas=al[:k], P00=p00[:k], t0=t0.value)

    def roundabout_lane(self, entrance_angle, exit_angle, distance) -> np.ndarray:
        pts = np.zeros((64, 2), np.float32)
        k = self.lib.ilqg_ref_roundabout_lane(C.c_float(entrance_angle), C.c_float(exit_angle),
                                              C.c_float(distance), _ptr(pts), 64)
        return pts[:k]

    def draw_circle(self, cx, cy, radius, num_segments) -> np.ndarray:
        pts = np.zeros((num_segments + 1, 2), np.float32)
        k = self.lib.ilqg_ref_draw_circle(C.c_float(cx), C.c_float(cy), C.c_float(radius),
                                          num_segments, _ptr(pts), num_segments + 1)
        return pts[:k]

    def polyline_constraint(self, pts, threshold: float, keep_left: bool, mu: float, lambda_step: float, xy):
        """The reference's own Polyline2SignedDistanceConstraint at the points `xy` [count][2]:
        rows of (g, lambda, d/dx, d/dy, d2/dx2, d2/dxdy, d2/dy2)."""
        pts = np.ascontiguousarray(pts, np.float32)
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros((len(xy), 7), np.float32)
        self.lib.ilqg_ref_polyline_constraint.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                                          C.c_void_p, C.c_int, C.c_void_p]
        rc = self.lib.ilqg_ref_polyline_constraint(_ptr(pts), len(pts), threshold, int(keep_left), mu, lambda_step,
                                                   _ptr(xy), len(xy), _ptr(out))
        assert rc == 0
        return out

    def _sum_udim_sq(self, which: int) -> int:
        return {INTERSECTION: 12, ROUNDABOUT: 16, AIR3D: 2, OVERTAKING: 12, COLLISION: 8, REACHABILITY2: 8, REACHABILITY3: 12, REACHABILITY1: 1, DUBINS_ORIGIN: 2, REACHABILITY_2P: 8, MODIFIED_AIR3D: 8, MODIFIED_INTERSECTION: 12, SKELETON: 8,
                INTERSECTION_REACHABILITY: 12}[which]
