"""solve() with max_iterations = 1..4: A/B/K of the LAST iteration and the traces, GPU vs oracle (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import grav_comp_guess, reference_set, standing_state
from mpc_ilqr_mujoco_b200 import gpu, Config
from oracle import pyoracle as po
tag = sys.argv[1] if len(sys.argv) > 1 else "standing"
policy = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = Config().build_weights()
win = reference_set(tag).window(0, 25)
x0 = standing_state(); ug = grav_comp_guess(x0)
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
for m in (1, 2, 3, 4):
    og = gpu.default_options(); og.max_iterations = m
    oo = po.default_options(); oo.max_iterations = m
    sg = gpu.H1IlqrBatch(w, N=25, batch=1, options=og); sg.set_kernel_policy(policy)
    so = po.OracleSolver(w, 25, batch=1, options=oo)
    sg.set_reference_window(*win, shared=True); so.set_reference_window(*win)
    so.initialize(x0, False, ug); sg.initialize(x0[None], None, ug)
    co = so.solve(x0); cg, it, st = sg.solve(x0[None])
    A, B = sg.get_linearization(); K, kff = sg.get_gains()
    ct, at = sg.solve_trace(); cto, ato = so.trace()
    print(m, "A", rel(A[0], so.get("A")), "B", rel(B[0], so.get("B")), "K", rel(K[0], so.get("K")), "kff", rel(kff[0], so.get("kff")))
    print("   gpu", ct[0][:m], at[0][:m].tolist(), "oracle", cto[:m], ato[:m].tolist())
