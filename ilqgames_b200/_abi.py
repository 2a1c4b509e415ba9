"""ctypes mirror of include/ilqg.h and a thin numpy-facing handle wrapper.

This module is binding glue only: it loads a shared library that implements
include/ilqg.h and moves numpy arrays across the C ABI.  The product library is
ilqgames_b200/lib/libilqg_b200.so (sm_100a CUDA); tests additionally load the
CPU oracle (oracle/_build/libilqg_oracle.so) through the SAME wrapper so parity
tests issue identical call sequences to both.  Nothing here falls back from one
library to the other.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

# ---- limits (include/ilqg.h) ------------------------------------------------
MAX_PLAYERS = 4
MAX_SUBSYSTEMS = 4
MAX_COSTS = 64
MAX_POLYLINES = 8
MAX_POLYLINE_POINTS = 128
MAX_PAIRS = 16
MAX_XDIM = 24
MAX_UDIM = 8
MAX_TIME_STEPS = 512

# ---- enums -------------------------------------------------------------------
DYN_NONE, DYN_CAR6D, DYN_UNICYCLE4D, DYN_AIR3D, DYN_CAR5D, DYN_DUBINS, DYN_TWO_PLAYER_UNICYCLE4D, DYN_POINT_MASS_2D = 0, 1, 2, 3, 4, 5, 6, 7
(COST_QUADRATIC, COST_QUADRATIC_POLYLINE2, COST_PROXIMITY, COST_SEMIQUADRATIC,
 COST_SEMIQUADRATIC_POLYLINE2, COST_POLYLINE2_SIGNED_DISTANCE, CONSTRAINT_PROXIMITY,
 CONSTRAINT_SINGLE_DIMENSION, COST_SIGNED_DISTANCE, COST_QUADRATIC_DIFFERENCE,
 CONSTRAINT_POLYLINE2_SIGNED_DISTANCE) = range(1, 12)
COST_SUM, COST_MAX, COST_MIN = 0, 1, 2
(STATUS_IDLE, STATUS_RUNNING, STATUS_CONVERGED, STATUS_MAX_ITERS,
 STATUS_LINESEARCH_FAILED, STATUS_NONFINITE) = range(6)

XS, US, PS, ALPHAS, LIN_A, LIN_B, QUAD_Q, QUAD_L, QUAD_R, QUAD_RGRAD, DELTA_XS = range(1, 12)
(STATUS, ITERS, MERIT, TOTAL_COSTS, LAMBDAS, MU, EXPECTED_DECREASE, STEP, BACKTRACKS,
 TIME_OF_EXTREME, X0, LQ_PS, LQ_ALPHAS, MAX_CONSTRAINT_ERROR, AL_SUCCESS, AL_ITERATES,
 AL_STATE, LQ_X0, WARM_XS, WARM_US, WARM_PS, WARM_ALPHAS) = range(12, 34)

OK = 0


class SubsystemDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("x_offset", C.c_int32), ("first_player", C.c_int32),
                ("reserved", C.c_int32), ("params", C.c_float * 4)]


class CostDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("player", C.c_int32), ("arg", C.c_int32),
                ("is_equality", C.c_int32), ("dim", C.c_int32 * 4), ("flag", C.c_int32),
                ("polyline", C.c_int32), ("weight", C.c_float), ("value", C.c_float),
                ("active_from", C.c_double), ("group", C.c_int32), ("group_is_min", C.c_int32)]


class ProblemDesc(C.Structure):
    _fields_ = [("num_time_steps", C.c_int32), ("num_players", C.c_int32), ("xdim", C.c_int32),
                ("udim", C.c_int32 * MAX_PLAYERS), ("time_step", C.c_double),
                ("initial_time", C.c_double), ("num_subsystems", C.c_int32),
                ("subsystems", SubsystemDesc * MAX_SUBSYSTEMS),
                ("state_regularization", C.c_float * MAX_PLAYERS),
                ("control_regularization", C.c_float * MAX_PLAYERS),
                ("cost_structure", C.c_int32 * MAX_PLAYERS), ("num_costs", C.c_int32),
                ("costs", CostDesc * MAX_COSTS), ("num_polylines", C.c_int32),
                ("polyline_start", C.c_int32 * (MAX_POLYLINES + 1)),
                ("polyline_points", (C.c_float * 2) * MAX_POLYLINE_POINTS)]


class SolverParams(C.Structure):
    """SolverParams, include/ilqgames/solver/solver_params.h:50-84 (defaults identical)."""
    _fields_ = [("convergence_tolerance", C.c_float), ("max_solver_iters", C.c_int32),
                ("linesearch", C.c_int32), ("initial_alpha_scaling", C.c_float),
                ("geometric_alpha_scaling", C.c_float), ("max_backtracking_steps", C.c_int32),
                ("expected_decrease_fraction", C.c_float), ("open_loop", C.c_int32),
                ("unconstrained_solver_max_iters", C.c_int32), ("geometric_mu_scaling", C.c_float),
                ("geometric_mu_downscaling", C.c_float),
                ("geometric_lambda_downscaling", C.c_float),
                ("constraint_error_tolerance", C.c_float), ("adaptive_regularization", C.c_int32),
                ("disable_convergence_exit", C.c_int32)]

    @classmethod
    def defaults(cls, **overrides) -> "SolverParams":
        p = cls(convergence_tolerance=1e-1, max_solver_iters=1000, linesearch=1,
                initial_alpha_scaling=0.5, geometric_alpha_scaling=0.5,
                max_backtracking_steps=10, expected_decrease_fraction=0.1, open_loop=0,
                unconstrained_solver_max_iters=10, geometric_mu_scaling=1.1,
                geometric_mu_downscaling=0.5, geometric_lambda_downscaling=0.5,
                constraint_error_tolerance=1e-1, adaptive_regularization=1,
                disable_convergence_exit=0)
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p


class Layout(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_time_steps", C.c_int32), ("num_players", C.c_int32),
                ("xdim", C.c_int32), ("total_udim", C.c_int32), ("udim", C.c_int32 * MAX_PLAYERS),
                ("u_offset", C.c_int32 * MAX_PLAYERS), ("num_pairs", C.c_int32),
                ("pair_player", C.c_int32 * MAX_PAIRS), ("pair_arg", C.c_int32 * MAX_PAIRS),
                ("pair_R_offset", C.c_int32 * MAX_PAIRS), ("pair_r_offset", C.c_int32 * MAX_PAIRS),
                ("R_floats", C.c_int32), ("r_floats", C.c_int32), ("num_constraints", C.c_int32),
                ("record_floats", C.c_int32), ("lambda_index", C.c_int32 * MAX_TIME_STEPS),
                ("compact_record_floats", C.c_int32)]


# every symbol include/ilqg.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "ilqg_create", "ilqg_create_multi", "ilqg_destroy", "ilqg_strerror", "ilqg_abi_struct_size", "ilqg_get_layout",
    "ilqg_upload_x0", "ilqg_upload_warmstart", "ilqg_upload", "ilqg_upload_lq",
    "ilqg_solve_begin", "ilqg_linearize_quadraticize", "ilqg_lq_backward", "ilqg_linesearch",
    "ilqg_iterate", "ilqg_al_update", "ilqg_overwrite_solution", "ilqg_al_post_solve",
    "ilqg_download", "ilqg_synchronize", "ilqg_kernel_launches", "ilqg_set_stream",
    "ilqg_profile", "ilqg_profile_read", "ilqg_reset", "ilqg_count_running",
    "ilqg_al_begin", "ilqg_al_advance", "ilqg_setup_next_receding_horizon", "ilqg_integrate_plan",
]

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.path.join(_REPO, "ilqgames_b200", "lib", "libilqg_b200.so")


class IlqgError(RuntimeError):
    pass


class Library:
    """A shared library implementing include/ilqg.h."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise IlqgError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; "
                "g.build()'). There is no CPU fallback for the product path.")
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L = self.lib
        vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.ilqg_create_multi.argtypes = [C.POINTER(ProblemDesc), C.POINTER(SolverParams), C.c_int,
                                        C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.ilqg_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(SolverParams), C.c_int, C.c_int,
                                  C.POINTER(vp)]
        L.ilqg_destroy.argtypes = [vp]
        L.ilqg_strerror.argtypes = [C.c_int]
        L.ilqg_strerror.restype = C.c_char_p
        L.ilqg_abi_struct_size.argtypes = [C.c_int]
        L.ilqg_abi_struct_size.restype = C.c_size_t
        L.ilqg_get_layout.argtypes = [vp, C.POINTER(Layout)]
        L.ilqg_upload_x0.argtypes = [vp, vp, C.c_size_t]
        L.ilqg_upload_warmstart.argtypes = [vp, vp, vp, vp, vp]
        L.ilqg_upload.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.ilqg_upload_lq.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        for name in ("ilqg_solve_begin", "ilqg_linearize_quadraticize", "ilqg_lq_backward",
                     "ilqg_linesearch", "ilqg_al_update", "ilqg_al_post_solve", "ilqg_synchronize"):
            getattr(L, name).argtypes = [vp]
        L.ilqg_iterate.argtypes = [vp, C.c_int, ip]
        L.ilqg_overwrite_solution.argtypes = [vp, C.c_int]
        L.ilqg_download.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.ilqg_kernel_launches.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.ilqg_set_stream.argtypes = [vp, vp]
        L.ilqg_reset.argtypes = [vp, C.c_int]
        L.ilqg_count_running.argtypes = [vp, ip]
        L.ilqg_al_begin.argtypes = [vp, C.c_int, C.c_float]
        L.ilqg_al_advance.argtypes = [vp, ip]
        L.ilqg_setup_next_receding_horizon.argtypes = [vp, vp, C.c_double, C.c_double,
                                                       C.POINTER(C.c_double)]
        L.ilqg_integrate_plan.argtypes = [vp, vp, C.c_double, C.c_double, vp]
        L.ilqg_profile.argtypes = [vp, C.c_int]
        L.ilqg_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_double),
                                        C.POINTER(C.c_longlong)]
        for name in ABI_SYMBOLS:
            if name not in ("ilqg_strerror", "ilqg_abi_struct_size"):
                getattr(L, name).restype = C.c_int

    def check(self, rc: int, what: str = ""):
        if rc != OK:
            msg = self.lib.ilqg_strerror(rc).decode()
            raise IlqgError(f"{what or 'ilqg call'} failed: {msg} (code {rc})")

    def verify_struct_sizes(self):
        for which, cls in enumerate((ProblemDesc, SolverParams, Layout, CostDesc, SubsystemDesc)):
            got = self.lib.ilqg_abi_struct_size(which)
            if got != C.sizeof(cls):
                raise IlqgError(f"ABI mismatch for {cls.__name__}: C={got} ctypes={C.sizeof(cls)}")


_product: Optional[Library] = None


def product_library() -> Library:
    """The sm_100a CUDA implementation.  Fails loudly when it has not been built."""
    global _product
    if _product is None:
        _product = Library(PRODUCT_LIB)
        _product.verify_struct_sizes()
    return _product


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Handle:
    """One ilqg_handle: `batch` independent games on one device."""

    def __init__(self, lib: Library, desc: ProblemDesc, params: SolverParams, batch: int,
                 device=0):
        """device: a CUDA ordinal, or a sequence of them (ilqg_create_multi: the batch sharded over
        several GPUs of the node behind this one handle)."""
        self.lib = lib
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*device)
            lib.check(lib.lib.ilqg_create_multi(C.byref(desc), C.byref(params), batch, devs, len(device),
                                                C.byref(self._h)), "ilqg_create_multi")
        else:
            lib.check(lib.lib.ilqg_create(C.byref(desc), C.byref(params), batch, device,
                                          C.byref(self._h)), "ilqg_create")
        self.layout = Layout()
        lib.check(lib.lib.ilqg_get_layout(self._h, C.byref(self.layout)), "ilqg_get_layout")
        lo = self.layout
        self.B, self.T, self.N, self.n, self.M = (lo.batch, lo.num_time_steps, lo.num_players,
                                                  lo.xdim, lo.total_udim)
        self._shapes = {
            XS: ((self.T, self.n), np.float32), US: ((self.T, self.M), np.float32),
            PS: ((self.T, self.M, self.n), np.float32), ALPHAS: ((self.T, self.M), np.float32),
            LQ_PS: ((self.T, self.M, self.n), np.float32),
            LQ_ALPHAS: ((self.T, self.M), np.float32),
            LIN_A: ((self.T, self.n, self.n), np.float32),
            LIN_B: ((self.T, self.n, self.M), np.float32),
            QUAD_Q: ((self.T, self.N, self.n, self.n), np.float32),
            QUAD_L: ((self.T, self.N, self.n), np.float32),
            QUAD_R: ((self.T, lo.R_floats), np.float32),
            QUAD_RGRAD: ((self.T, lo.r_floats), np.float32),
            DELTA_XS: ((self.T, self.n), np.float32), STATUS: ((), np.int32),
            ITERS: ((), np.int32), MERIT: ((), np.float32),
            TOTAL_COSTS: ((self.N,), np.float32),
            LAMBDAS: ((lo.num_constraints, self.T), np.float32), MU: ((), np.float32),
            EXPECTED_DECREASE: ((), np.float32), STEP: ((), np.float32),
            BACKTRACKS: ((), np.int32), TIME_OF_EXTREME: ((self.N,), np.int32),
            X0: ((self.n,), np.float32), MAX_CONSTRAINT_ERROR: ((), np.float32),
            AL_SUCCESS: ((), np.int32), AL_ITERATES: ((), np.int32), AL_STATE: ((), np.int32),
            LQ_X0: ((self.n,), np.float32),
            WARM_XS: ((self.T, self.n), np.float32), WARM_US: ((self.T, self.M), np.float32),
            WARM_PS: ((self.T, self.M, self.n), np.float32),
            WARM_ALPHAS: ((self.T, self.M), np.float32),
        }

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.lib.ilqg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data movement ----------------------------------------------------------
    def shape(self, what: int):
        shp, dt = self._shapes[what]
        return (self.B,) + shp, dt

    def upload_x0(self, x0):
        a = _f32(x0)
        assert a.shape == (self.B, self.n), a.shape
        self.lib.check(self.lib.lib.ilqg_upload_x0(self._h, a.ctypes.data, a.nbytes), "upload_x0")

    def upload_warmstart(self, xs=None, us=None, Ps=None, alphas=None):
        arrs = []
        for a, what in ((xs, XS), (us, US), (Ps, PS), (alphas, ALPHAS)):
            if a is None:
                arrs.append(None)
            else:
                a = _f32(a)
                assert a.shape == self.shape(what)[0], (a.shape, self.shape(what)[0])
                arrs.append(a)
        ptrs = [a.ctypes.data if a is not None else None for a in arrs]
        self.lib.check(self.lib.lib.ilqg_upload_warmstart(self._h, *ptrs), "upload_warmstart")

    def upload(self, what: int, arr):
        shp, dt = self.shape(what)
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.shape == shp, (a.shape, shp)
        self.lib.check(self.lib.lib.ilqg_upload(self._h, what, a.ctypes.data, a.nbytes), "upload")

    def upload_lq(self, A, Bs, Q, l, R, r):
        arrs = [_f32(a) for a in (A, Bs, Q, l, R, r)]
        for a, what in zip(arrs, (LIN_A, LIN_B, QUAD_Q, QUAD_L, QUAD_R, QUAD_RGRAD)):
            assert a.shape == self.shape(what)[0], (what, a.shape, self.shape(what)[0])
        self.lib.check(self.lib.lib.ilqg_upload_lq(self._h, *[a.ctypes.data for a in arrs]),
                       "upload_lq")

    def download(self, what: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        shp, dt = self.shape(what)
        if out is None:
            out = np.empty(shp, dtype=dt)
        assert out.shape == shp and out.dtype == dt and out.flags.c_contiguous
        self.lib.check(self.lib.lib.ilqg_download(self._h, what, out.ctypes.data, out.nbytes),
                       f"download({what})")
        return out

    def upload_x0_ptr(self, ptr: int, nbytes: int):
        self.lib.check(self.lib.lib.ilqg_upload_x0(self._h, C.c_void_p(ptr), nbytes), "upload_x0")

    def download_ptr(self, what: int, ptr: int, nbytes: int):
        self.lib.check(self.lib.lib.ilqg_download(self._h, what, C.c_void_p(ptr), nbytes),
                       f"download({what})")

    # -- hot path ---------------------------------------------------------------
    def solve_begin(self):
        self.lib.check(self.lib.lib.ilqg_solve_begin(self._h), "solve_begin")

    def linearize_quadraticize(self):
        self.lib.check(self.lib.lib.ilqg_linearize_quadraticize(self._h), "linearize_quadraticize")

    def lq_backward(self):
        self.lib.check(self.lib.lib.ilqg_lq_backward(self._h), "lq_backward")

    def linesearch(self):
        self.lib.check(self.lib.lib.ilqg_linesearch(self._h), "linesearch")

    def iterate(self, max_iters: int, want_count: bool = False) -> Optional[int]:
        if want_count:
            done = C.c_int(0)
            self.lib.check(self.lib.lib.ilqg_iterate(self._h, max_iters, C.byref(done)), "iterate")
            return done.value
        self.lib.check(self.lib.lib.ilqg_iterate(self._h, max_iters, None), "iterate")
        return None

    def count_running(self) -> int:
        out = C.c_int(0)
        self.lib.check(self.lib.lib.ilqg_count_running(self._h, C.byref(out)), "count_running")
        return out.value

    def solve(self, max_passes: int = 10_000, chunk: int = 4) -> int:
        """ILQSolver::Solve after solve_begin(): passes until no instance is RUNNING.  Returns the
        number of passes issued."""
        done = 0
        while done < max_passes:
            self.iterate(chunk)
            done += chunk
            if self.count_running() == 0:
                break
        return done

    def al_begin(self, max_iterates: int, constraint_error_tolerance: float):
        self.lib.check(self.lib.lib.ilqg_al_begin(self._h, int(max_iterates),
                                                  float(constraint_error_tolerance)), "al_begin")

    def al_advance(self) -> int:
        out = C.c_int(0)
        self.lib.check(self.lib.lib.ilqg_al_advance(self._h, C.byref(out)), "al_advance")
        return out.value

    def al_update(self):
        self.lib.check(self.lib.lib.ilqg_al_update(self._h), "al_update")

    def al_post_solve(self):
        self.lib.check(self.lib.lib.ilqg_al_post_solve(self._h), "al_post_solve")

    def overwrite_solution(self, only_successful: bool = False):
        self.lib.check(self.lib.lib.ilqg_overwrite_solution(self._h, int(only_successful)),
                       "overwrite_solution")

    def setup_next_receding_horizon(self, x0, t0: float, planner_runtime: float) -> float:
        """Problem::SetUpNextRecedingHorizon for every game; returns the new OperatingPoint::t0."""
        a = _f32(x0)
        assert a.shape == (self.B, self.n), a.shape
        new_t0 = C.c_double(0.0)
        self.lib.check(self.lib.lib.ilqg_setup_next_receding_horizon(
            self._h, a.ctypes.data, float(t0), float(planner_runtime), C.byref(new_t0)),
            "setup_next_receding_horizon")
        return new_t0.value

    def integrate_plan(self, x0, t0: float, t: float) -> np.ndarray:
        """MultiPlayerIntegrableSystem::Integrate(t0, t, x0, plan) for every game: [B][n] -> [B][n]."""
        a = _f32(x0)
        assert a.shape == (self.B, self.n), a.shape
        out = np.zeros_like(a)
        self.lib.check(self.lib.lib.ilqg_integrate_plan(self._h, a.ctypes.data, float(t0), float(t),
                                                        out.ctypes.data), "integrate_plan")
        return out

    def synchronize(self):
        self.lib.check(self.lib.lib.ilqg_synchronize(self._h), "synchronize")

    RESET_SOLVER, RESET_MULTIPLIERS, RESET_SOLUTION, RESET_LAMBDAS, RESET_MU = 1, 2, 4, 8, 16

    def reset(self, mask: int = 1):
        self.lib.check(self.lib.lib.ilqg_reset(self._h, mask), "reset")

    def set_stream(self, cuda_stream: int):
        """cuda_stream: a cudaStream_t as an integer (e.g. torch.cuda.current_stream().cuda_stream)."""
        self.lib.check(self.lib.lib.ilqg_set_stream(self._h, C.c_void_p(cuda_stream)), "set_stream")

    def profile(self, enable: bool):
        self.lib.check(self.lib.lib.ilqg_profile(self._h, int(enable)), "profile")

    KERNEL_NAMES = ("linearize_quadraticize", "lq_backward", "linesearch", "solve_begin",
                    "ls_eval_fresh", "ls_eval_queued", "ls_decide", "ls_eval_begin")

    def profile_read(self):
        """{kernel name: (total_ms, launches)} since profile(True)."""
        out = {}
        for k, name in enumerate(self.KERNEL_NAMES):
            ms, n = C.c_double(0), C.c_longlong(0)
            self.lib.check(self.lib.lib.ilqg_profile_read(self._h, k, C.byref(ms), C.byref(n)),
                           "profile_read")
            out[name] = (ms.value, n.value)
        return out

    def kernel_launches(self) -> int:
        out = C.c_longlong(0)
        self.lib.check(self.lib.lib.ilqg_kernel_launches(self._h, C.byref(out)), "kernel_launches")
        return out.value
