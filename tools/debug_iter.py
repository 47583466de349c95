"""Stage-by-stage comparison GPU vs oracle over iLQR iterations (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import grav_comp_guess, make_oracle, standing_state
from mpc_ilqr_mujoco_b200 import gpu
tag = sys.argv[1] if len(sys.argv) > 1 else "standing"
policy = int(sys.argv[2]) if len(sys.argv) > 2 else 2
so, w, win = make_oracle(tag, N=25, t0=0, linearization=0)
sg = gpu.H1IlqrBatch(w, N=25, batch=1)
sg.set_kernel_policy(policy)
sg.set_reference_window(*win, shared=True)
x0 = standing_state(); ug = grav_comp_guess(x0)
so.initialize(x0, False, ug); sg.initialize(x0[None], None, ug)
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
lam = 1e-6
for it in range(5):
    so.set_lambda(lam); sg.set_regularization(lam)
    so.rollout_nominal(x0); sg.rollout_nominal(x0[None])
    xg, ugp = sg.get_trajectory()
    print(it, "rollout x", rel(xg[0], so.get("xbar")))
    so.linearize(); sg.linearize()
    A, B = sg.get_linearization()
    print(it, "A", rel(A[0], so.get("A")), "B", rel(B[0], so.get("B")))
    so.cost_quadratics(); sg.cost_quadratics()
    so.backward_pass(); sg.backward_pass()
    K, kff = sg.get_gains()
    print(it, "K", rel(K[0], so.get("K")), "kff", rel(kff[0], so.get("kff")))
    ok_o, c_o, a_o = so.line_search(x0)
    ok, c, a = sg.line_search(x0[None])
    print(it, "ls oracle", ok_o, c_o, a_o, "gpu", bool(ok[0]), c[0], a[0])
    if ok_o: lam = max(lam / 2, 1e-6)
