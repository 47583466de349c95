"""CPU tests of the CUDA kernels' warp-cooperative phase functions through the lane-emulation build
(tests/emul): same source as the kernels (csrc/h1_dyn.cuh, csrc/h1_costq.cuh), lanes run sequentially. Checks the
index tables, sparse factorisation, tangent solve and Hessian assembly against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, make_oracle

EMUL = os.path.join(ROOT, "tests", "emul")
P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
IP = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-C", EMUL], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(EMUL, "libdyn_emul.so")), C.CDLL(os.path.join(EMUL, "libcostq_emul.so"))


def states(n, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 51))
    x[:, :3] = rng.uniform(-1, 1, (n, 3)); x[:, 0] += 4.0; x[:, 2] = 1.0 + rng.uniform(-0.05, 0.1, n)
    q = rng.normal(size=(n, 4)) * 0.1; q[:, 0] += 1
    x[:, 3:7] = q
    x[:, 7:26] = rng.uniform(-0.5, 0.5, (n, 19)); x[:, 26:] = rng.uniform(-1, 1, (n, 25))
    return x, rng.uniform(-50, 50, (n, 19))


def test_lane_parallel_dynamics_matches_dense_oracle(emul, oracle):
    x, u = states(64, 1)
    xe = np.empty_like(x); com = np.empty((64, 3))
    assert emul[0].emul_dyn_step(64, P(x), P(u), P(xe), P(com)) == 0
    assert np.abs(xe - oracle.dyn_step(x, u)).max() < 2e-12
    assert max(np.abs(com[i] - oracle.dyn_com(x[i])).max() for i in range(64)) < 1e-14


def test_thread_sequential_dynamics_matches_dense_oracle(emul, oracle):
    """csrc/h1_dyn_seq.cuh (one thread per f_D: single DFS walk, contact folded into the foot bodies' spatial
    inertia / wrench, compact in-place L'DL) against the dense oracle, incl. clamped torques, feet in deep contact
    and in the air; its factor equals the warp-cooperative one (what the linearization kernels consume)."""
    x, u = states(96, 7)
    u[::5, 2] = 500.0; u[1::5, 7] = -500.0
    x[::4, 2] -= 0.08          # soles well inside the ground
    x[1::4, 2] += 0.3          # airborne
    xe = np.empty_like(x); com = np.empty((96, 3))
    nf = 25 * 11 + 50
    fs = np.zeros((96, nf)); fw = np.zeros((96, nf))
    assert emul[0].emul_dyn_step_seq(96, P(x), P(u), P(xe), P(com), P(fs), P(fw)) == 0
    assert np.abs(xe - oracle.dyn_step(x, u)).max() < 2e-12
    assert max(np.abs(com[i] - oracle.dyn_com(x[i])).max() for i in range(96)) < 1e-14
    assert np.abs(fs - fw).max() <= 1e-11 * np.abs(fw).max()
    com2 = np.empty((96, 3))
    assert emul[0].emul_dyn_com_seq(96, P(x), P(com2)) == 0
    assert np.abs(com2 - com).max() < 1e-14


def test_tangent_linearization_matches_oracle_ad(emul, oracle):
    x, u = states(12, 2)
    u[::3, 3] = 400.0
    for i in range(12):
        A, B = oracle.dyn_linearize_ad(x[i], u[i])
        Ae = np.empty((51, 51), order="F"); Be = np.empty((51, 19), order="F")
        assert emul[0].emul_dyn_linearize_analytic(P(x[i]), P(u[i]), P(Ae), P(Be)) == 0
        assert np.abs(Ae - A).max() <= 1e-10 * np.abs(A).max()
        assert np.abs(Be - B).max() <= 1e-10 * np.abs(B).max()


def test_direction_per_thread_linearization_matches_oracle_ad(emul, oracle):
    """csrc/h1_lin_dirs.cuh (one thread per column, sequential DFS walk with cumulative wrenches) against the
    oracle's forward-mode AD through its dense f_D: 1e-10 relative, incl. clamped torques, raw quaternions."""
    x, u = states(12, 4)
    u[::3, 3] = 400.0
    for i in range(12):
        A, B = oracle.dyn_linearize_ad(x[i], u[i])
        Ae = np.empty((51, 51), order="F"); Be = np.empty((51, 19), order="F")
        assert emul[0].emul_dyn_linearize_dirs(P(x[i]), P(u[i]), P(Ae), P(Be)) == 0
        assert np.abs(Ae - A).max() <= 1e-10 * np.abs(A).max()
        assert np.abs(Be - B).max() <= 1e-10 * np.abs(B).max()


def test_sparse_column_linearization_matches_oracle_ad(emul, oracle):
    """the per-direction tangent functions of k_linearize_tangents (joint directions walk only subtree(joint) with dual numbers, the rigid base directions use the contact-only / rotation-covariance tangents) composed into columns,
    the x / y columns are unit vectors; against the oracle's forward-mode AD, incl. feet in contact and clamps."""
    x, u = states(12, 9)
    u[::3, 3] = 400.0
    x[::2, 2] -= 0.06
    for i in range(12):
        A, B = oracle.dyn_linearize_ad(x[i], u[i])
        Ae = np.empty((51, 51), order="F"); Be = np.empty((51, 19), order="F")
        assert emul[0].emul_dyn_linearize_cols(P(x[i]), P(u[i]), P(Ae), P(Be)) == 0
        assert np.abs(Ae - A).max() <= 1e-10 * np.abs(A).max()
        assert np.abs(Be - B).max() <= 1e-10 * np.abs(B).max()
        assert (Ae[:, 0] == np.eye(51)[0]).all() and (Ae[:, 1] == np.eye(51)[1]).all()


@pytest.mark.parametrize("tag", ["standing", "walking"])
def test_cost_quadratics_phases_match_oracle(emul, oracle, tag):
    s, w, win = make_oracle(tag)
    x_ref, u_ref, com_ref, ee_ref, stance, cv = win
    rng = np.random.default_rng(3)
    xb = x_ref + rng.normal(size=x_ref.shape) * 0.05
    xb[:, 7:26] += rng.uniform(-0.6, 0.6, (26, 19))
    ub = rng.uniform(-250, 250, (25, 19))
    s.set("xbar", xb); s.set("ubar", ub); s.cost_quadratics()
    lx, lu, lxx, luu = s.get("lx"), s.get("lu"), s.get("lxx"), s.get("luu")
    for t in range(26):
        glx, glu = np.zeros(51), np.zeros(19)
        glxx, gluu = np.zeros((51, 51)), np.zeros((19, 19))
        u = ub[t] if t < 25 else np.zeros(19)
        ur = u_ref[t] if t < 25 else np.zeros(19)
        rc = emul[1].emul_cost_quadratics(C.byref(w), P(xb[t]), P(u), P(np.ascontiguousarray(x_ref[t])), P(ur),
                                          P(np.ascontiguousarray(com_ref[t])), P(np.ascontiguousarray(cv[t])),
                                          P(np.ascontiguousarray(ee_ref[t])), IP(np.ascontiguousarray(stance[t])),
                                          int(t == 25), P(glx), P(glu), P(glxx), P(gluu))
        assert rc == 0
        assert np.abs(glx - lx[t]).max() <= 1e-12 * np.abs(lx[t]).max()
        assert np.abs(glxx - lxx[t]).max() <= 1e-12 * np.abs(lxx[t]).max()
        if t < 25:
            assert np.abs(glu - lu[t]).max() <= 1e-12 * max(np.abs(lu[t]).max(), 1e-300)
            assert np.abs(gluu - luu[t]).max() <= 1e-12 * np.abs(luu[t]).max()


def test_quad_cooperative_dynamics_matches_dense_oracle(emul, oracle):
    """csrc/h1_dyn_quad.cuh (four lanes per f_D, one per kinematic chain: articulated-body recursion with the parent
    pose recovered by the inverse joint transform, torso shared by the two arm lanes, 6 x 6 base block summed over the
    lanes) against the dense oracle; the four lanes run as four host threads, the kernel's xor-shuffles as exchanges
    between barriers. Includes clamped torques, un-normalised quaternions, feet deep in contact and in the air."""
    lib = C.CDLL(os.path.join(EMUL, "libquad_emul.so"))
    x, u = states(96, 9)
    u[::5, 2] = 500.0; u[1::5, 7] = -500.0
    x[::4, 2] -= 0.08
    x[1::4, 2] += 0.3
    x[::7, 3:7] *= 1.07
    xe = np.zeros_like(x); com = np.empty((96, 3)); com2 = np.empty((96, 3))
    assert lib.emul_dyn_step_quad(96, P(x), P(u), P(xe), P(com), P(com2)) == 0
    assert np.abs(xe - oracle.dyn_step(x, u)).max() < 5e-12
    assert max(np.abs(com[i] - oracle.dyn_com(x[i])).max() for i in range(96)) < 1e-14
    assert np.abs(com2 - com).max() < 1e-14
    # zero torques (u = nullptr) = the plant's free fall
    xz = np.zeros_like(x)
    assert lib.emul_dyn_step_quad(96, P(x), None, P(xz), P(com), P(com2)) == 0
    assert np.abs(xz - oracle.dyn_step(x, np.zeros_like(u))).max() < 5e-12
