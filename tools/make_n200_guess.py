"""Generates tests/golden/n200_guess.npz: a dynamically consistent 200-step trajectory (x0, U[200][19]) that stays
upright, used as the initial guess of the N = 200 parity test (BASELINE config 4). It is the closed loop of the
oracle's own N = 25 MPC on the STANDING reference (plant = the oracle's f_D), started from the standing pose: the
constant gravity-compensation guess the reference uses for a cold start lets the robot fall within 4 s at N = 200
(the balance cost then takes the square root of a negative CoM height -> NaN, the `diverging` case of the same test),
and the walking closed loop itself loses balance after ~140 steps, so a convergent long-horizon walking solve needs a
guess that keeps the robot on its feet for the whole horizon.
Run from the repo root: python tools/make_n200_guess.py   (about 2-3 minutes of CPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import grav_comp_guess, make_oracle, po, reference_set, standing_state  # noqa: E402

STEPS = 200
so, w, _ = make_oracle("standing")
refs = reference_set("standing")
x = standing_state()
ug = grav_comp_guess(standing_state())
X = [x.copy()]
U = []
for k in range(STEPS):
    so.set_reference_window(*refs.window(min(k, refs.T - 26), 25))
    u, c = so.mpc_step(x, ug)
    x = po.dyn_step(x, u)[0]
    U.append(u.copy()); X.append(x.copy())
    if k % 20 == 0:
        print(k, "cost %.2f" % c, "z %.3f" % x[2], flush=True)
X = np.array(X); U = np.array(U)
assert np.isfinite(X).all()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "n200_guess.npz"), x0=X[0], U=U, X=X)
print("saved; z range", X[:, 2].min(), X[:, 2].max())
