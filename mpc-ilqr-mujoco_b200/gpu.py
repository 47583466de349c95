"""ctypes binding of libh1ilqr.so — the C ABI declared in include/h1ilqr.h.

There is no fallback: importing this module without the built CUDA library, or creating a batch without a
B200-class device, raises. Arrays are numpy float64 / int32, C-contiguous, instance-major; matrices are
returned as [instance][knot][col][row] (column-major per knot, Eigen's layout).
"""
import ctypes as C
import os

import numpy as np

from .ctypes_defs import (H1Model, H1SolverOptions, H1StageTimes, H1Weights, NQ, NU, NV, NX, c_double_p, c_int_p,
                          dptr, iptr)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("H1ILQR_LIB") or os.path.join(_HERE, "lib", "libh1ilqr.so")   # (override: A/B experiment builds)

EXPORTS = [
    "h1ilqr_default_options", "h1ilqr_create", "h1ilqr_destroy", "h1ilqr_last_error", "h1ilqr_batch",
    "h1ilqr_horizon", "h1ilqr_set_weights", "h1ilqr_set_weight_matrices", "h1ilqr_set_reference_window", "h1ilqr_initialize", "h1ilqr_solve",
    "h1ilqr_mpc_step", "h1ilqr_mpc_reset", "h1ilqr_rollout_nominal", "h1ilqr_linearize", "h1ilqr_cost_quadratics",
    "h1ilqr_backward_pass", "h1ilqr_line_search", "h1ilqr_total_cost", "h1ilqr_dynamics_step", "h1ilqr_bias_forces",
    "h1ilqr_reference_kinematics", "h1ilqr_reference_com_velocity", "h1ilqr_reference_ee_velocity", "h1ilqr_linearize_state",
    "h1ilqr_limit_penalties", "h1ilqr_stage_cost", "h1ilqr_get_status", "h1ilqr_sole_points", "h1ilqr_set_trajectory", "h1ilqr_get_trajectory", "h1ilqr_get_gains",
    "h1ilqr_set_gains", "h1ilqr_get_linearization", "h1ilqr_set_linearization", "h1ilqr_get_cost_quadratics",
    "h1ilqr_set_cost_quadratics", "h1ilqr_set_previous_solution", "h1ilqr_get_previous_solution", "h1ilqr_get_regularization", "h1ilqr_set_regularization", "h1ilqr_get_solve_trace",
    "h1ilqr_upload_inputs", "h1ilqr_host_register", "h1ilqr_host_unregister", "h1ilqr_run_resident_steps", "h1ilqr_time_stage", "h1ilqr_set_reference_table", "h1ilqr_run_closed_loop", "h1ilqr_measure_fp64_peak", "h1ilqr_measure_fp64_mma_peak", "h1ilqr_enable_stage_timing", "h1ilqr_set_kernel_policy", "h1ilqr_get_stage_times", "h1ilqr_stream", "h1_default_dynamics_model",
    "h1_default_cost_model",
]

KERNELS_AUTO, KERNELS_COOPERATIVE, KERNELS_BATCHED = 0, 1, 2

_lib = None


def lib():
    """Load libh1ilqr.so (raises if the CUDA extension has not been built: no CPU path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(make -C mpc-ilqr-mujoco_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.h1ilqr_last_error.restype = C.c_char_p
        L.h1ilqr_stream.restype = C.c_void_p
        L.h1_default_dynamics_model.restype = C.POINTER(H1Model)
        L.h1_default_cost_model.restype = C.POINTER(H1Model)
        _lib = L
    return _lib


class H1IlqrError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise H1IlqrError(f"h1ilqr error {rc}: {lib().h1ilqr_last_error().decode()}")


def default_options():
    o = H1SolverOptions()
    lib().h1ilqr_default_options(C.byref(o))
    return o


def default_dynamics_model():
    m = H1Model()
    C.memmove(C.byref(m), lib().h1_default_dynamics_model(), C.sizeof(H1Model))
    return m


def default_cost_model():
    m = H1Model()
    C.memmove(C.byref(m), lib().h1_default_cost_model(), C.sizeof(H1Model))
    return m


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, tuple(shape))
    return a


class H1IlqrBatch:
    """`batch` independent H1 MPC/iLQR instances resident on one GPU (one C-ABI handle)."""

    def __init__(self, weights, N=25, batch=1, device=0, options=None, dyn_model=None, cost_model=None):
        self.N, self.B = int(N), int(batch)
        self.opt = options if options is not None else default_options()
        self._h = C.c_void_p()
        self._pinned = []
        _check(lib().h1ilqr_create(C.byref(dyn_model) if dyn_model is not None else None,
                                   C.byref(cost_model) if cost_model is not None else None, C.byref(self.opt),
                                   C.c_int(self.B), C.c_int(self.N), C.c_int(device), C.byref(self._h)))
        self.set_weights(weights)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.unpin_all()
            lib().h1ilqr_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- configuration ----
    def set_weights(self, w: H1Weights):
        _check(lib().h1ilqr_set_weights(self._h, C.byref(w)))

    def set_weight_matrices(self, Q=None, R=None, Qf=None):
        """Full symmetric Q [51,51], R [19,19], Qf [51,51] (ilqr.cpp:145-150); all None = diagonal weights again."""
        cm = lambda M, n: np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(n, n).T) if M is not None else None
        q, r, f = cm(Q, NX), cm(R, NU), cm(Qf, NX)     # column-major copies, alive across the call
        _check(lib().h1ilqr_set_weight_matrices(self._h, dptr(q), dptr(r), dptr(f)))

    def set_reference_window(self, x_ref, u_ref, com_ref, ee_ref, stance, com_vel_ref=None, shared=True):
        n = 1 if shared else self.B
        N, N1 = self.N, self.N + 1
        x_ref = _f(x_ref).reshape(n, N1, NX)
        u_ref = _f(u_ref).reshape(n, N, NU)
        com_ref = _f(com_ref).reshape(n, N1, 3)
        ee_ref = _f(ee_ref).reshape(n, N1, 2, 3)
        stance = np.ascontiguousarray(stance, dtype=np.int32).reshape(n, N1, 2)
        cv = _f(com_vel_ref).reshape(n, N1, 3) if com_vel_ref is not None else None
        _check(lib().h1ilqr_set_reference_window(self._h, dptr(x_ref), dptr(u_ref), dptr(com_ref), dptr(ee_ref),
                                                 iptr(stance), dptr(cv), C.c_int(int(shared))))

    def pin_host(self, *arrays):
        """Page-lock caller-owned numpy arrays that are handed to set_reference_window / mpc_step every step, so their
        host <-> device copies are direct DMA transfers. Returns the (contiguous) arrays to keep and pass on."""
        out = []
        for a in arrays:
            a = np.ascontiguousarray(a)
            _check(lib().h1ilqr_host_register(self._h, C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))
            self._pinned.append(a)
            out.append(a)
        return out if len(out) != 1 else out[0]

    def unpin_all(self):
        for a in self._pinned:
            lib().h1ilqr_host_unregister(self._h, C.c_void_p(a.ctypes.data))
        self._pinned = []

    # ---- solver ----
    def initialize(self, x0, warm=None, u_init=None):
        x0 = _f(x0).reshape(self.B, NX)
        wp = np.ascontiguousarray(warm, dtype=np.int32).reshape(self.B) if warm is not None else None
        shared = 1
        if u_init is not None:
            u_init = _f(u_init)
            shared = int(u_init.size == NU)
        _check(lib().h1ilqr_initialize(self._h, dptr(x0), iptr(wp), dptr(u_init), C.c_int(shared)))

    def solve(self, x0):
        x0 = _f(x0).reshape(self.B, NX)
        cost = np.empty(self.B)
        iters = np.empty(self.B, dtype=np.int32)
        status = np.empty(self.B, dtype=np.int32)
        rc = lib().h1ilqr_solve(self._h, dptr(x0), dptr(cost), iptr(iters), iptr(status))
        if rc not in (0, -3):
            _check(rc)
        return cost, iters, status

    def mpc_step(self, x_measured, u_init=None):
        x = _f(x_measured).reshape(self.B, NX)
        shared = 1
        if u_init is not None:
            u_init = _f(u_init)
            shared = int(u_init.size == NU)
        ua = np.empty((self.B, NU))
        cost = np.empty(self.B)
        rc = lib().h1ilqr_mpc_step(self._h, dptr(x), dptr(u_init), C.c_int(shared), dptr(ua), dptr(cost))
        self.last_rc = rc
        if rc not in (0, -3):   # -3 = H1ILQR_ENOTFINITE: outputs are delivered, get_status() tells which instance
            _check(rc)
        return ua, cost

    def get_status(self):
        """(status[B] (0 ok, 1 non-finite cost / gains), iters[B]) of the last solve / MPC step."""
        st = np.empty(self.B, dtype=np.int32)
        it = np.empty(self.B, dtype=np.int32)
        _check(lib().h1ilqr_get_status(self._h, iptr(st), iptr(it)))
        return st, it

    def mpc_reset(self):
        _check(lib().h1ilqr_mpc_reset(self._h))

    # ---- stages ----
    def rollout_nominal(self, x0):
        _check(lib().h1ilqr_rollout_nominal(self._h, dptr(_f(x0).reshape(self.B, NX))))

    def linearize(self):
        _check(lib().h1ilqr_linearize(self._h))

    def cost_quadratics(self):
        _check(lib().h1ilqr_cost_quadratics(self._h))

    def backward_pass(self):
        _check(lib().h1ilqr_backward_pass(self._h))

    def line_search(self, x0):
        ok = np.empty(self.B, dtype=np.int32)
        nc = np.empty(self.B)
        ai = np.empty(self.B, dtype=np.int32)
        _check(lib().h1ilqr_line_search(self._h, dptr(_f(x0).reshape(self.B, NX)), iptr(ok), dptr(nc), iptr(ai)))
        return ok, nc, ai

    def total_cost(self):
        c = np.empty(self.B)
        _check(lib().h1ilqr_total_cost(self._h, dptr(c)))
        return c

    def dynamics_step(self, x, u):
        x = _f(x).reshape(-1, NX)
        u = _f(u).reshape(-1, NU)
        xn = np.empty_like(x)
        _check(lib().h1ilqr_dynamics_step(self._h, C.c_int(x.shape[0]), dptr(x), dptr(u), dptr(xn)))
        return xn

    def bias_forces(self, x):
        x = _f(x).reshape(-1, NX)
        b = np.empty((x.shape[0], NV))
        _check(lib().h1ilqr_bias_forces(self._h, C.c_int(x.shape[0]), dptr(x), dptr(b)))
        return b

    def reference_kinematics(self, x):
        x = _f(x).reshape(-1, NX)
        com = np.empty((x.shape[0], 3))
        ee = np.empty((x.shape[0], 2, 3))
        _check(lib().h1ilqr_reference_kinematics(self._h, C.c_int(x.shape[0]), dptr(x), dptr(com), dptr(ee)))
        return com, ee

    def reference_com_velocity(self, x):
        """Whole-body CoM velocity [n][3] (world frame) on the dynamics model (robot_utils.cpp:388-397)."""
        x = _f(x).reshape(-1, NX)
        cv = np.empty((x.shape[0], 3))
        _check(lib().h1ilqr_reference_com_velocity(self._h, C.c_int(x.shape[0]), dptr(x), dptr(cv)))
        return cv

    def reference_ee_velocity(self, x):
        """World velocity [n][2][3] of the two ankle-body origins (robot_utils.cpp:405-412)."""
        x = _f(x).reshape(-1, NX)
        ev = np.empty((x.shape[0], 2, 3))
        _check(lib().h1ilqr_reference_ee_velocity(self._h, C.c_int(x.shape[0]), dptr(x), dptr(ev)))
        return ev

    def linearize_state(self, x, u, mode=0, eps=1e-5):
        """(A [51][51], B [51][19]) of one (x, u) pair, row-major numpy views of the column-major C arrays transposed:
        A[i, j] = d x_next_i / d x_j. mode 1 = RobotUtils::linearizeDynamicsFD (forward differences, eps)."""
        A = np.empty((NX, NX)); B = np.empty((NU, NX))
        _check(lib().h1ilqr_linearize_state(self._h, C.c_int(mode), C.c_double(eps), dptr(_f(x).reshape(NX)), dptr(_f(u).reshape(NU)),
                                            dptr(A), dptr(B)))
        return A.T, B.T

    def limit_penalties(self, x, u=None):
        """constraintCost / Gradients / Hessian diagonals of n (x, u) pairs (robot_utils.cpp:615-778)."""
        x = _f(x).reshape(-1, NX); n = x.shape[0]
        uu = _f(u).reshape(n, NU) if u is not None else None
        c = np.empty(n); gx = np.empty((n, NX)); gu = np.empty((n, NU)); hx = np.empty((n, NX)); hu = np.empty((n, NU))
        _check(lib().h1ilqr_limit_penalties(self._h, C.c_int(n), dptr(x), dptr(uu), dptr(c), dptr(gx), dptr(gu), dptr(hx), dptr(hu)))
        return c, gx, gu, hx, hu

    def stage_cost(self, x, u, x_ref, u_ref=None, com_ref=None):
        """RobotUtils::stageCost (u given) / terminalCost (u None) of n states against reference rows."""
        x = _f(x).reshape(-1, NX); n = x.shape[0]
        uu = _f(u).reshape(n, NU) if u is not None else None
        ur = _f(u_ref).reshape(n, NU) if u_ref is not None else None
        cr = _f(com_ref).reshape(n, 3) if com_ref is not None else None
        c = np.empty(n)
        _check(lib().h1ilqr_stage_cost(self._h, C.c_int(n), dptr(x), dptr(uu), dptr(_f(x_ref).reshape(n, NX)), dptr(ur), dptr(cr), dptr(c)))
        return c

    def sole_points(self, x):
        """World positions [n][8][3] of the sole contact points (left foot's four first)."""
        x = _f(x).reshape(-1, NX)
        pts = np.empty((x.shape[0], 8, 3))
        _check(lib().h1ilqr_sole_points(self._h, C.c_int(x.shape[0]), dptr(x), dptr(pts)))
        return pts

    # ---- accessors ----
    def get_trajectory(self):
        xb = np.empty((self.B, self.N + 1, NX))
        ub = np.empty((self.B, self.N, NU))
        _check(lib().h1ilqr_get_trajectory(self._h, dptr(xb), dptr(ub)))
        return xb, ub

    def set_trajectory(self, xbar=None, ubar=None):
        xb = _f(xbar, (self.B, self.N + 1, NX)) if xbar is not None else None
        ub = _f(ubar, (self.B, self.N, NU)) if ubar is not None else None
        _check(lib().h1ilqr_set_trajectory(self._h, dptr(xb), dptr(ub)))

    def get_gains(self):
        K = np.empty((self.B, self.N, NX, NU))
        k = np.empty((self.B, self.N, NU))
        _check(lib().h1ilqr_get_gains(self._h, dptr(K), dptr(k)))
        return K, k

    def set_gains(self, K, kff):
        _check(lib().h1ilqr_set_gains(self._h, dptr(_f(K, (self.B, self.N, NX, NU))), dptr(_f(kff, (self.B, self.N, NU)))))

    def get_linearization(self):
        A = np.empty((self.B, self.N, NX, NX))
        Bm = np.empty((self.B, self.N, NU, NX))
        _check(lib().h1ilqr_get_linearization(self._h, dptr(A), dptr(Bm)))
        return A, Bm

    def set_linearization(self, A, Bm):
        _check(lib().h1ilqr_set_linearization(self._h, dptr(_f(A, (self.B, self.N, NX, NX))),
                                              dptr(_f(Bm, (self.B, self.N, NU, NX)))))

    def get_cost_quadratics(self):
        lx = np.empty((self.B, self.N + 1, NX))
        lu = np.empty((self.B, self.N, NU))
        lxx = np.empty((self.B, self.N + 1, NX, NX))
        luu = np.empty((self.B, self.N, NU, NU))
        _check(lib().h1ilqr_get_cost_quadratics(self._h, dptr(lx), dptr(lu), dptr(lxx), dptr(luu)))
        return lx, lu, lxx, luu

    def set_cost_quadratics(self, lx, lu, lxx, luu):
        _check(lib().h1ilqr_set_cost_quadratics(self._h, dptr(_f(lx)), dptr(_f(lu)), dptr(_f(lxx)), dptr(_f(luu))))

    def set_previous_solution(self, prev_xbar, prev_ubar):
        """Hand over the previous MPC solution a warm start shifts (MPC::prev_xbar_ / prev_ubar_)."""
        _check(lib().h1ilqr_set_previous_solution(self._h, dptr(_f(prev_xbar, (self.B, self.N + 1, NX))),
                                                  dptr(_f(prev_ubar, (self.B, self.N, NU)))))

    def get_previous_solution(self):
        xb = np.empty((self.B, self.N + 1, NX))
        ub = np.empty((self.B, self.N, NU))
        _check(lib().h1ilqr_get_previous_solution(self._h, dptr(xb), dptr(ub)))
        return xb, ub

    def get_regularization(self):
        lam = np.empty(self.B)
        _check(lib().h1ilqr_get_regularization(self._h, dptr(lam)))
        return lam

    def set_regularization(self, lam):
        lam = _f(np.atleast_1d(lam))
        _check(lib().h1ilqr_set_regularization(self._h, dptr(lam), C.c_int(int(lam.size == 1))))

    def solve_trace(self):
        ct = np.empty((self.B, self.opt.max_iterations))
        at = np.empty((self.B, self.opt.max_iterations, 2), dtype=np.int32)
        _check(lib().h1ilqr_get_solve_trace(self._h, dptr(ct), iptr(at)))
        return ct, at

    def upload_inputs(self, x_measured, u_init=None):
        x = _f(x_measured).reshape(self.B, NX)
        shared = 1
        if u_init is not None:
            u_init = _f(u_init)
            shared = int(u_init.size == NU)
        _check(lib().h1ilqr_upload_inputs(self._h, dptr(x), dptr(u_init), C.c_int(shared)))

    def run_resident_steps(self, steps, cold_each_step=True, graph=False):
        """`steps` MPC steps with inputs already on the device; returns CUDA-event milliseconds on the handle's stream.
        graph: capture the step once into a CUDA graph and replay it."""
        ms = C.c_double()
        _check(lib().h1ilqr_run_resident_steps(self._h, C.c_int(steps), C.c_int(int(bool(cold_each_step)) | (2 if graph else 0)),
                                               C.byref(ms)))
        return ms.value

    def set_reference_table(self, refs, schedule_offset=False):
        """Upload a ReferenceSet's full tables for the device-resident closed loop."""
        x = _f(refs.x_ref_full); com = _f(refs.com_ref_full); ee = _f(refs.ee_pos_ref_full).reshape(-1, 6)
        cv = _f(refs.com_vel_ref_full)
        ct = np.ascontiguousarray(refs.contact, dtype=np.int32).reshape(-1, 2)
        _check(lib().h1ilqr_set_reference_table(self._h, C.c_int(x.shape[0]), dptr(x), dptr(com), dptr(ee), dptr(cv),
                                                C.c_int(ct.shape[0]), iptr(ct), C.c_int(int(schedule_offset))))

    def run_closed_loop(self, steps, t_idx0=None, x_start=None, u_init=None, graph=True, logs=True):
        """`steps` closed-loop MPC steps of every instance on the device (plant = f_D). Returns a dict with x_final,
        cost / iters / u logs (when `logs`) and the CUDA-event milliseconds of the whole loop."""
        t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t_idx0, dtype=np.int32), (self.B,))) if t_idx0 is not None else None
        xs = _f(x_start).reshape(self.B, NX) if x_start is not None else None
        shared = 1
        if u_init is not None:
            u_init = _f(u_init)
            shared = int(u_init.size == NU)
        xf = np.empty((self.B, NX))
        cl = np.empty((steps, self.B)) if logs else None
        il = np.empty((steps, self.B), dtype=np.int32) if logs else None
        ul = np.empty((steps, self.B, NU)) if logs else None
        ms = C.c_double()
        rc = lib().h1ilqr_run_closed_loop(self._h, C.c_int(steps), iptr(t0), dptr(xs), dptr(u_init), C.c_int(shared),
                                          C.c_int(int(graph)), dptr(xf), dptr(cl), iptr(il), dptr(ul), C.byref(ms))
        if rc not in (0, -3):
            _check(rc)
        return {"x_final": xf, "cost": cl, "iters": il, "u": ul, "ms": ms.value, "rc": rc}

    STAGES = {"factor": 0, "linearize": 1, "cost_quadratics": 2, "backward": 3, "line_search": 4}

    def time_stage(self, stage, reps=1):
        """Device milliseconds PER LAUNCH of one stage on the current trajectory / derivatives / gains."""
        ms = C.c_double()
        _check(lib().h1ilqr_time_stage(self._h, C.c_int(self.STAGES[stage]), C.c_int(reps), C.byref(ms)))
        return ms.value / reps

    def measure_fp64_peak(self):
        t = C.c_double()
        _check(lib().h1ilqr_measure_fp64_peak(self._h, C.byref(t)))
        return t.value

    def measure_fp64_mma_peak(self):
        t = C.c_double()
        _check(lib().h1ilqr_measure_fp64_mma_peak(self._h, C.byref(t)))
        return t.value

    def set_kernel_policy(self, policy):
        """KERNELS_AUTO (0), KERNELS_COOPERATIVE (1: warp per unit, latency) or KERNELS_BATCHED (2: thread per unit)."""
        _check(lib().h1ilqr_set_kernel_policy(self._h, C.c_int(int(policy))))

    def enable_stage_timing(self, flag=True):
        _check(lib().h1ilqr_enable_stage_timing(self._h, C.c_int(int(flag))))

    def stage_times(self):
        t = H1StageTimes()
        _check(lib().h1ilqr_get_stage_times(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in H1StageTimes._fields_}

    def stream(self):
        return lib().h1ilqr_stream(self._h)
