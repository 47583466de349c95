"""ctypes mirrors of the plain-C structs in include/h1_model.h and include/h1ilqr.h."""
import ctypes as C

NB, NQ, NV, NX, NU, NFOOT, NCP, NALPHA = 20, 26, 25, 51, 19, 2, 4, 8
LIN_ANALYTIC, LIN_FD = 0, 1
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class H1Model(C.Structure):
    _fields_ = [
        ("parent", C.c_int * NB), ("axis", C.c_int * NB), ("has_rfix", C.c_int * NB),
        ("pos", (C.c_double * 3) * NB), ("rfix", (C.c_double * 9) * NB), ("mass", C.c_double * NB),
        ("ipos", (C.c_double * 3) * NB), ("inertia", (C.c_double * 6) * NB),
        ("armature", C.c_double * NV), ("damping", C.c_double * NV),
        ("jnt_range", (C.c_double * 2) * NU), ("ctrl_range", (C.c_double * 2) * NU),
        ("foot_body", C.c_int * NFOOT), ("foot_pts", ((C.c_double * 3) * NCP) * NFOOT),
        ("gravity", C.c_double * 3), ("timestep", C.c_double),
        ("contact_kn", C.c_double), ("contact_bn", C.c_double), ("contact_bt", C.c_double),
        ("contact_eps", C.c_double), ("total_mass", C.c_double),
    ]


class H1Weights(C.Structure):
    _fields_ = [
        ("Qdiag", C.c_double * NX), ("Rdiag", C.c_double * NU), ("Qfdiag", C.c_double * NX),
        ("w_com", C.c_double), ("w_com_vel", C.c_double), ("w_ee_pos", C.c_double), ("w_ee_vel", C.c_double),
        ("w_upright", C.c_double), ("w_balance", C.c_double),
        ("w_joint_limits", C.c_double), ("w_control_limits", C.c_double),
    ]


class H1SolverOptions(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int), ("tolerance", C.c_double), ("reg_init", C.c_double),
        ("reg_min", C.c_double), ("reg_max", C.c_double), ("accept_margin", C.c_double),
        ("fd_eps", C.c_double), ("divergence_cost", C.c_double), ("linearization", C.c_int),
        ("alphas", C.c_double * NALPHA),
    ]


class H1StageTimes(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double), ("rollout_ms", C.c_double), ("linearize_ms", C.c_double),
        ("cost_quadratics_ms", C.c_double), ("backward_ms", C.c_double), ("line_search_ms", C.c_double),
        ("launches", C.c_int),
    ]


def dptr(a):
    """numpy float64 C-contiguous array -> double*"""
    return a.ctypes.data_as(c_double_p) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None
