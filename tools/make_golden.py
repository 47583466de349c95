#!/usr/bin/env python3
"""Generate tests/golden/oracle_golden.npz: seeded inputs and the CPU oracle's outputs for every stage of the path.

The reference ships no golden vectors, so these pin the ORACLE (regression) and give the GPU
tests fixed targets that do not depend on rebuilding the oracle. Run from the repo root in the build container.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import grav_comp_guess, make_oracle, standing_state  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

out = {}
rng = np.random.default_rng(2024)
n = 32
x = np.zeros((n, 51))
x[:, :3] = rng.uniform(-1, 1, (n, 3)); x[:, 2] = 1.02 + rng.uniform(-0.03, 0.08, n)
q = rng.normal(size=(n, 4)) * 0.1; q[:, 0] += 1
x[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
x[:, 7:26] = rng.uniform(-0.5, 0.5, (n, 19)); x[:, 26:] = rng.uniform(-1, 1, (n, 25))
u = rng.uniform(-60, 60, (n, 19)); u[::5] *= 8.0
out["dyn_x"], out["dyn_u"], out["dyn_xnext"] = x, u, po.dyn_step(x, u)
out["dyn_bias"] = np.array([po.dyn_bias(r) for r in x])
out["dyn_com"] = np.array([po.dyn_com(r) for r in x])
A, B = zip(*(po.dyn_linearize_ad(x[i], u[i]) for i in range(4)))
out["lin_A"], out["lin_B"] = np.array([np.asarray(a).T for a in A]), np.array([np.asarray(b).T for b in B])  # [i][col][row]

for tag in ("standing", "walking"):
    s, w, win = make_oracle(tag)
    x0 = standing_state(); ug = grav_comp_guess(x0)
    s.initialize(x0, False, ug)
    s.rollout_nominal(x0); s.linearize(); s.cost_quadratics(); s.backward_pass()
    for k in ("xbar", "ubar", "A", "B", "lx", "lu", "lxx", "luu", "K", "kff"):
        out[f"{tag}_iter0_{k}"] = s.get(k)
    ok, nc, ai = s.line_search(x0)
    out[f"{tag}_iter0_ls"] = np.array([float(ok), nc, float(ai)])
    s.mpc_reset(); s.initialize(x0, False, ug)
    c = s.solve(x0)
    ct, at = s.trace()
    out[f"{tag}_solve_cost"], out[f"{tag}_solve_trace"], out[f"{tag}_solve_alpha"] = np.array([c]), ct, at
    out[f"{tag}_solve_xbar"], out[f"{tag}_solve_ubar"] = s.get("xbar"), s.get("ubar")
    out[f"{tag}_u_guess"] = ug
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **out)
print("wrote", len(out), "arrays")
