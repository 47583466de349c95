"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_build/liboracle.so (the CPU oracle).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from mpc_ilqr_mujoco_b200.ctypes_defs import (H1Model, H1SolverOptions, H1Weights, NALPHA, NU, NX, c_double_p,
                                              c_int_p, dptr, iptr)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_default_dynamics_model.restype = C.POINTER(H1Model)
        L.orc_default_cost_model.restype = C.POINTER(H1Model)
        L.orc_cost_term.restype = C.c_double
        L.orc_limit_cost.restype = C.c_double
        L.orc_create.restype = C.c_void_p
        L.orc_total_cost.restype = C.c_double
        L.orc_get_lambda.restype = C.c_double
        L.orc_mpc_step_batch.restype = C.c_double
        L.orc_set_lambda.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _lib = L
    return _lib


def default_options():
    o = H1SolverOptions()
    lib().orc_default_options(C.byref(o))
    return o


def dynamics_model():
    m = H1Model()
    C.memmove(C.byref(m), lib().orc_default_dynamics_model(), C.sizeof(H1Model))
    return m


def cost_model():
    m = H1Model()
    C.memmove(C.byref(m), lib().orc_default_cost_model(), C.sizeof(H1Model))
    return m


def _mp(m):
    return C.byref(m) if m is not None else None


def dyn_step(x, u, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, NX)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, NU)
    xn = np.empty_like(x)
    lib().orc_dyn_step(_mp(model), C.c_int(x.shape[0]), dptr(x), dptr(u), dptr(xn))
    return xn


def dyn_linearize(x, u, eps=1e-5, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    A = np.empty((NX, NX), order="F")
    B = np.empty((NX, NU), order="F")
    lib().orc_dyn_linearize(_mp(model), dptr(x), dptr(u), C.c_double(eps), A.ctypes.data_as(c_double_p),
                            B.ctypes.data_as(c_double_p))
    return A, B


def dyn_linearize_ad(x, u, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    A = np.empty((NX, NX), order="F")
    B = np.empty((NX, NU), order="F")
    lib().orc_dyn_linearize_ad(_mp(model), dptr(x), dptr(u), A.ctypes.data_as(c_double_p), B.ctypes.data_as(c_double_p))
    return A, B


def dyn_com(x, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(3)
    lib().orc_dyn_com(_mp(model), dptr(x), dptr(out))
    return out


def dyn_com_vel(x, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(3)
    lib().orc_dyn_com_vel(_mp(model), dptr(x), dptr(out))
    return out


def dyn_bias(x, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(25)
    lib().orc_dyn_bias(_mp(model), dptr(x), dptr(out))
    return out


def dyn_body_pos(x, body, model=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(3)
    lib().orc_dyn_body_pos(_mp(model), dptr(x), C.c_int(body), dptr(out))
    return out


TERM_COM, TERM_COM_VEL, TERM_EE_POS, TERM_EE_VEL, TERM_UPRIGHT, TERM_BALANCE = range(6)


def cost_term(term, x, target, w, ee=0, mode=0, model=None):
    """mode 0: value; 1: AD; 2: analytic. Returns (value, g[51], H[51,51])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    t = np.zeros(3)
    if target is not None:
        t[:len(target)] = target
    g = np.zeros(NX)
    H = np.zeros((NX, NX), order="F")
    v = lib().orc_cost_term(_mp(model), C.c_int(term), C.c_int(ee), dptr(x), dptr(t), C.c_double(w), C.c_int(mode),
                            dptr(g), H.ctypes.data_as(c_double_p))
    return v, g, H


class OracleSolver:
    """`batch` independent oracle iLQR instances sharing one problem (weights, reference window)."""

    def __init__(self, weights, N, batch=1, options=None, dyn_model=None, cost_model=None):
        self.N, self.batch = N, batch
        self.opt = options if options is not None else default_options()
        self.h = C.c_void_p(lib().orc_create(_mp(dyn_model), _mp(cost_model), C.byref(weights), C.byref(self.opt),
                                             C.c_int(batch), C.c_int(N)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def set_weight_matrices(self, Q=None, R=None, Qf=None):
        """Whole cost matrices Q [51,51], R [19,19], Qf [51,51] (ilqr.cpp:145-150); all None = diagonal weights."""
        if Q is None:
            lib().orc_set_weight_matrices(self.h, None, None, None)
            return
        cm = lambda M, n: np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(n, n).T)
        q, r, f = cm(Q, NX), cm(R, NU), cm(Qf, NX)
        lib().orc_set_weight_matrices(self.h, dptr(q), dptr(r), dptr(f))

    def use_ad(self, flag):
        lib().orc_use_ad(self.h, C.c_int(int(flag)))

    def set_reference_window(self, x_ref, u_ref, com_ref, ee_ref, stance, com_vel_ref=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x_ref, u_ref, com_ref, ee_ref = f(x_ref), f(u_ref), f(com_ref), f(ee_ref)
        stance = np.ascontiguousarray(stance, dtype=np.int32)
        cv = f(com_vel_ref) if com_vel_ref is not None else None
        lib().orc_set_reference_window(self.h, dptr(x_ref), dptr(u_ref), dptr(com_ref), dptr(ee_ref), iptr(stance),
                                       dptr(cv))

    def initialize(self, x0, warm=False, u_init=None, i=0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        u_init = np.zeros(NU) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
        lib().orc_initialize(self.h, C.c_int(i), dptr(x0), C.c_int(int(warm)), dptr(u_init))

    def rollout_nominal(self, x0, i=0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        lib().orc_rollout_nominal(self.h, C.c_int(i), dptr(x0))

    def linearize(self, i=0):
        lib().orc_linearize(self.h, C.c_int(i))

    def cost_quadratics(self, i=0):
        lib().orc_cost_quadratics(self.h, C.c_int(i))

    def backward_pass(self, i=0):
        lib().orc_backward_pass(self.h, C.c_int(i))

    def line_search(self, x0, i=0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        nc, ai = C.c_double(), C.c_int()
        ok = lib().orc_line_search(self.h, C.c_int(i), dptr(x0), C.byref(nc), C.byref(ai))
        return bool(ok), nc.value, ai.value

    def total_cost(self, i=0):
        return lib().orc_total_cost(self.h, C.c_int(i))

    def solve(self, x0, i=0):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        c = C.c_double()
        lib().orc_solve(self.h, C.c_int(i), dptr(x0), C.byref(c))
        return c.value

    def mpc_step(self, x, u_init=None, i=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u_init = np.zeros(NU) if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
        ua = np.empty(NU)
        c = C.c_double()
        lib().orc_mpc_step(self.h, C.c_int(i), dptr(x), dptr(u_init), dptr(ua), C.byref(c))
        return ua, c.value

    def mpc_reset(self, i=0):
        lib().orc_mpc_reset(self.h, C.c_int(i))

    def mpc_step_batch(self, x0, u_init, threads):
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(self.batch, NX)
        u_init = np.ascontiguousarray(u_init, dtype=np.float64)
        ua = np.empty((self.batch, NU))
        cost = np.empty(self.batch)
        secs = lib().orc_mpc_step_batch(self.h, dptr(x0), dptr(u_init), dptr(ua), dptr(cost), C.c_int(threads))
        return secs, ua, cost

    def iters(self, i=0):
        return lib().orc_iters(self.h, C.c_int(i))

    def get_lambda(self, i=0):
        return lib().orc_get_lambda(self.h, C.c_int(i))

    def set_lambda(self, lam, i=0):
        lib().orc_set_lambda(self.h, C.c_int(i), C.c_double(lam))

    def trace(self, i=0):
        ct = np.empty(self.opt.max_iterations)
        at = np.empty((self.opt.max_iterations, 2), dtype=np.int32)
        lib().orc_get_trace(self.h, C.c_int(i), dptr(ct), iptr(at))
        return ct, at

    def margins(self, i=0):
        """Distance (cost units) of every accept / reject / stop decision of the last solve from flipping:
        (ls_margin [max_iterations][2], stop_margin [max_iterations]); -1 = decision not taken."""
        lm = np.empty((self.opt.max_iterations, 2))
        sm = np.empty(self.opt.max_iterations)
        lib().orc_get_margins(self.h, C.c_int(i), dptr(lm), dptr(sm))
        return lm, sm

    _shapes = {
        "xbar": lambda N: (N + 1, NX), "ubar": lambda N: (N, NU), "K": lambda N: (N, NX, NU),
        "kff": lambda N: (N, NU), "A": lambda N: (N, NX, NX), "B": lambda N: (N, NU, NX),
        "lx": lambda N: (N + 1, NX), "lu": lambda N: (N, NU), "lxx": lambda N: (N + 1, NX, NX),
        "luu": lambda N: (N, NU, NU),
    }

    def get(self, name, i=0):
        """Raw buffers. Matrices come back as [knot][col][row] (column-major per knot):
        use .transpose(0, 2, 1) to index them [knot][row][col]."""
        out = np.empty(self._shapes[name](self.N))
        getattr(lib(), "orc_get_" + name)(self.h, C.c_int(i), dptr(out))
        return out

    def set(self, name, arr, i=0):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._shapes[name](self.N), (arr.shape, self._shapes[name](self.N))
        getattr(lib(), "orc_set_" + name)(self.h, C.c_int(i), dptr(arr))
