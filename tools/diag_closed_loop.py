"""Diagnostic (GPU box): stage-by-stage GPU-vs-oracle comparison along the hard end-of-table walking closed loop
(tests/test_gpu_workloads.py::test_config2_closed_loop segment b). Prints, per MPC step, the relative differences of
the warm-start guess, A/B, cost quadratics, gains and line-search result computed on IDENTICAL inputs, next to the
difference of the full MPC steps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import grav_comp_guess, make_oracle, po, reference_set, rel_err, standing_state  # noqa: E402
from mpc_ilqr_mujoco_b200 import gpu  # noqa: E402

T0 = int(sys.argv[1]) if len(sys.argv) > 1 else 360
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 16
so, w, _ = make_oracle("walking")
s2, _, _ = make_oracle("walking")          # mirror of `so` used for the hand-run iteration
sg = gpu.H1IlqrBatch(w, N=25, batch=1)
sh = gpu.H1IlqrBatch(w, N=25, batch=1)     # hand-run stages on the oracle's inputs
refs = reference_set("walking")
x = refs.x_ref_full[T0].copy()
ug = grav_comp_guess(standing_state())
for k in range(STEPS):
    win = refs.window(T0 + k, 25)
    for s in (so, s2):
        s.set_reference_window(*win)
    sg.set_reference_window(*win, shared=True); sh.set_reference_window(*win, shared=True)
    # hand-run first iteration on identical inputs (oracle's warm-start guess and lambda)
    lam = s2.get_lambda()
    s2.initialize(x, k > 0, ug); s2.rollout_nominal(x); s2.linearize(); s2.cost_quadratics(); s2.backward_pass()
    sh.set_trajectory(xbar=s2.get("xbar")[None], ubar=s2.get("ubar")[None])
    sh.rollout_nominal(x[None]); xr, _ = sh.get_trajectory()
    sh.set_trajectory(xbar=s2.get("xbar")[None], ubar=s2.get("ubar")[None])
    sh.linearize(); sh.cost_quadratics(); sh.set_regularization(lam); sh.backward_pass()
    A, B = sh.get_linearization(); lx, lu, lxx, luu = sh.get_cost_quadratics(); K, kff = sh.get_gains()
    dA = max(rel_err(A[0, t], s2.get("A")[t]) for t in range(25)); dB = max(rel_err(B[0, t], s2.get("B")[t]) for t in range(25))
    dl = max(rel_err(lxx[0, t], s2.get("lxx")[t]) for t in range(26)); dg = max(rel_err(lx[0, t], s2.get("lx")[t]) for t in range(26))
    dK = rel_err(K[0], s2.get("K")); dk = rel_err(kff[0], s2.get("kff"))
    dKt = [rel_err(K[0, t], s2.get("K")[t]) for t in range(25)]
    # gains from the oracle -> line search on both
    sh.set_gains(s2.get("K")[None], s2.get("kff")[None])
    ok_g, c_g, a_g = sh.line_search(x[None])
    ok_o, c_o, a_o = s2.line_search(x)
    print(f"step {k:2d} roll {rel_err(xr[0], s2.get('xbar')):.1e} dA {dA:.1e} dB {dB:.1e} dlx {dg:.1e} dlxx {dl:.1e} dK {dK:.1e} (worst knot {int(np.argmax(dKt))}: {max(dKt):.1e}) dk {dk:.1e} "
          f"maxK {np.abs(s2.get('K')).max():.0f} | LS same gains: a {a_g[0]}/{a_o} dc {abs(c_g[0]-c_o)/abs(c_o):.1e}", flush=True)
    # the real MPC steps
    uo, co = so.mpc_step(x, ug)
    ugp, cg = sg.mpc_step(x[None], ug)
    ct, at = sg.solve_trace(); cto, ato = so.trace(); it = so.iters()
    xg, ugt = sg.get_trajectory()
    print(f"        mpc: cost {co:.6f} d {abs(cg[0]-co)/abs(co):.1e} alpha eq {bool((at[0]==ato).all())} {ato[:it].tolist()} dx {rel_err(xg[0], so.get('xbar')):.1e} du {np.abs(ugt[0]-so.get('ubar')).max()/max(np.abs(so.get('ubar')).max(),1):.1e} du_apply {np.abs(ugp[0]-uo).max():.1e}", flush=True)
    # keep the mirror in lockstep with `so` (previous solution + lambda): undo the hand-run iteration's lambda, step it
    s2.set_lambda(lam)
    s2.mpc_step(x, ug)
    x = po.dyn_step(x, uo)[0]
