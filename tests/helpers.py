"""Shared test helpers: problem setup on the oracle side (CPU) used by both CPU and GPU tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mpc_ilqr_mujoco_b200 import Config  # noqa: E402
from mpc_ilqr_mujoco_b200.references import ReferenceSet, standing_state  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def oracle_kinematics(x):
    x = np.atleast_2d(x)
    com = np.array([po.dyn_com(r) for r in x])
    ee = np.array([[po.dyn_body_pos(r, 5), po.dyn_body_pos(r, 10)] for r in x])
    return com, ee


_REFS = {}


def reference_set(tag, kinematics=oracle_kinematics):
    key = (tag, kinematics)
    if key not in _REFS:
        d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
        _REFS[key] = ReferenceSet(d[f"{tag}_q"], d[f"{tag}_v"], d[f"{tag}_contact"], kinematics)
    return _REFS[key]


def grav_comp_guess(x0):
    """Cold-start guess of the reference (computeGravComp, robot_utils.cpp:844-866) incl. quirk Q15:
    torque i is read from qfrc_bias[7+i]; i = 18 reads one past the end, defined here as 0."""
    b = po.dyn_bias(x0)
    u = np.zeros(19)
    u[:18] = b[7:25]
    return u


def make_oracle(tag="standing", N=25, t0=0, batch=1, cfg=None, linearization=0):
    cfg = cfg or Config()
    w = cfg.build_weights()
    opt = po.default_options()
    opt.linearization = linearization  # 0 analytic (default), 1 forward differences (the reference's method)
    s = po.OracleSolver(w, N, batch=batch, options=opt)
    refs = reference_set(tag)
    win = refs.window(t0, N)
    s.set_reference_window(*win)
    return s, w, win


# ---- GPU-vs-oracle comparison of a full iLQR solve, with explicit near-tie accounting --------------------------------
# iLQR::solve takes discrete decisions (accept the FIRST alpha with cost < baseline - 1e-6, stop when |dcost| < 1e-4).
# Two fp64 implementations that agree to rounding can take different branches when a decision sits closer to its
# threshold than their rounding difference. Such instances are REPORTED as near-ties (the oracle
# records how far every decision was from flipping, OracleSolver.margins) and compared up to the fork only; they are
# never absorbed by a looser tolerance. TIE_REL is the north star's own tolerance on the per-iteration cost (1e-6
# relative): a decision closer to its threshold than that cannot be told apart at the accuracy the path claims.
TIE_REL = 1e-6
TOL = 1e-6


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def compare_solve(so, cost_g, iters_g, ct_g, at_g, x_g, u_g, label="", tol_xu=TOL):
    """Compare one instance's GPU solve (cost, iters, cost trace, alpha trace [it][2], xbar, ubar) with the oracle
    solver `so` that has just solved the same problem. Returns "match" or "near_tie"; raises on a real mismatch."""
    ct_o, at_o = so.trace()
    lm, sm = so.margins()
    it_o = so.iters()
    n = max(int(iters_g), int(it_o))
    fork = None
    for it in range(n):
        for a in range(2):
            if at_g[it][a] != at_o[it][a]:
                # a line search that one side ran and the other did not: the previous stop decision differed
                ran_o, ran_g = at_o[it][a] != -2, at_g[it][a] != -2
                if ran_o and ran_g:
                    margin = lm[it][a]
                elif a == 0 and it > 0:
                    margin = sm[it - 1] if sm[it - 1] >= 0 else min([m for m in lm[it - 1] if m >= 0] or [-1.0])
                else:
                    margin = lm[it][0]
                fork = (it, a, margin)
                break
        if fork:
            break
    if fork is None and int(iters_g) != int(it_o):
        fork = (min(int(iters_g), int(it_o)) - 1, 2, sm[min(int(iters_g), int(it_o)) - 1])
    if fork is None:
        assert rel_err(ct_g[:it_o], ct_o[:it_o]) < TOL, (label, "cost trace", ct_g[:it_o], ct_o[:it_o])
        co = ct_o[it_o - 1] if it_o > 0 else None
        assert co is None or abs(cost_g - co) <= TOL * abs(co), (label, "cost", cost_g, co)
        assert rel_err(x_g, so.get("xbar")) < tol_xu, (label, "xbar", rel_err(x_g, so.get("xbar")))
        assert np.abs(u_g - so.get("ubar")).max() <= tol_xu * max(np.abs(so.get("ubar")).max(), 1.0), (label, "ubar")
        return "match"
    it, a, margin = fork
    scale = max(1.0, abs(ct_o[it]) if ct_o[it] != 0.0 else abs(ct_o[max(it - 1, 0)]))
    assert 0.0 <= margin <= TIE_REL * scale, (
        f"{label}: decisions differ at iteration {it} (attempt {a}) although the oracle's decision margin there is "
        f"{margin:.3e} (> {TIE_REL:.0e} x cost {scale:.3e}): GPU {at_g[:n].tolist()} oracle {at_o[:n].tolist()}")
    if it > 0:   # everything before the fork must still agree
        assert rel_err(ct_g[:it], ct_o[:it]) < TOL, (label, "cost trace before the fork")
    return "near_tie"


def oracle_solves(weights, wins, x0s, u_guess, N=25, threads=None, options=None):
    """Cold-start oracle solves of independent instances (own window each), run on a thread pool (the C++ side
    releases the GIL). Returns the list of OracleSolver objects, each holding its solution and traces."""
    import os as _os
    from concurrent.futures import ThreadPoolExecutor
    solvers = []
    for i in range(len(x0s)):
        s = po.OracleSolver(weights, N, batch=1, options=options)
        s.set_reference_window(*(a[i] for a in wins))
        solvers.append(s)

    def run(i):
        solvers[i].initialize(x0s[i], False, u_guess)
        solvers[i].solve(x0s[i])
    with ThreadPoolExecutor(max_workers=threads or _os.cpu_count() or 4) as ex:
        list(ex.map(run, range(len(x0s))))
    return solvers
