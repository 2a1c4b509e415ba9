"""ctypes access to the oracle-only probe functions (not part of include/ilqg.h)."""
import ctypes as C

import numpy as np


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def segment_closest_point(lib, a, b, q):
    closest = np.zeros(2, np.float32)
    is_end, ssd = C.c_int(), C.c_float()
    rc = lib.lib.ilqg_oracle_segment_closest_point(
        C.c_float(a[0]), C.c_float(a[1]), C.c_float(b[0]), C.c_float(b[1]), C.c_float(q[0]),
        C.c_float(q[1]), _fp(closest), C.byref(is_end), C.byref(ssd))
    return rc, closest, bool(is_end.value), ssd.value


def polyline_closest_point(lib, handle, p, q):
    closest = np.zeros(2, np.float32)
    is_vertex, seg, is_end, ssd = C.c_int(), C.c_int(), C.c_int(), C.c_float()
    rc = lib.lib.ilqg_oracle_polyline_closest_point(
        handle._h, p, C.c_float(q[0]), C.c_float(q[1]), _fp(closest), C.byref(is_vertex),
        C.byref(seg), C.byref(ssd), C.byref(is_end))
    assert rc == 0
    return closest, bool(is_vertex.value), seg.value, ssd.value, bool(is_end.value)


def evaluate_record(lib, handle, c, x, lam=0.0, mu=0.0, with_al=False):
    x = np.ascontiguousarray(x, np.float32)
    out = C.c_float()
    rc = lib.lib.ilqg_oracle_evaluate_record(handle._h, c, _fp(x), len(x), C.c_float(lam),
                                             C.c_float(mu), int(with_al), C.byref(out))
    assert rc == 0
    return out.value


def quadraticize_record(lib, handle, c, x, lam=0.0, mu=0.0):
    x = np.ascontiguousarray(x, np.float32)
    n = len(x)
    hess = np.zeros((n, n), np.float32)
    grad = np.zeros(n, np.float32)
    rc = lib.lib.ilqg_oracle_quadraticize_record(handle._h, c, _fp(x), n, C.c_float(lam),
                                                 C.c_float(mu), _fp(hess), _fp(grad))
    assert rc == 0
    return hess, grad


def dynamics(lib, handle, x, u):
    x = np.ascontiguousarray(x, np.float32)
    u = np.ascontiguousarray(u, np.float32)
    xdot = np.zeros_like(x)
    xnext = np.zeros_like(x)
    rc = lib.lib.ilqg_oracle_dynamics(handle._h, _fp(x), _fp(u), _fp(xdot), _fp(xnext))
    assert rc == 0
    return xdot, xnext
