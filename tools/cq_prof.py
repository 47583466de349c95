"""Per-phase cycles of k_cost_quadratics from a -DCQ_PROF build (GPU box): for the two warps of a knot pair, work before / wait at
the pair barrier that ends each phase, averaged over all knots.
Build first (here): cd mpc-ilqr-mujoco_b200/csrc && nvcc $(Makefile flags) -DCQ_PROF -shared -o ../lib/libh1ilqr_prof.so h1ilqr_capi.cu model_tables.cpp -lcudart
usage: H1ILQR_LIB=$PWD/mpc-ilqr-mujoco_b200/lib/libh1ilqr_prof.so python tools/cq_prof.py [B]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_ilqr_mujoco_b200 import Config, gpu  # noqa: E402
from mpc_ilqr_mujoco_b200 import workloads as wl  # noqa: E402
from mpc_ilqr_mujoco_b200.references import standing_state  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = 25
s = gpu.H1IlqrBatch(Config().build_weights(), N=N, batch=B)
win, x0, _ = wl.walking_instances(np.arange(B), s.reference_kinematics)
s.set_reference_window(*win, shared=False)
ug = np.zeros(19); ug[:18] = s.bias_forces(standing_state()[None])[0][7:25]
s.initialize(x0, None, ug)
s.rollout_nominal(x0); s.linearize(); s.cost_quadratics(); s.backward_pass()
L = ctypes.CDLL(gpu.LIB_PATH)
buf = (ctypes.c_ulonglong * 64)()
L.h1ilqr_debug_cq_prof(buf)
ms = s.time_stage("cost_quadratics", 1)
L.h1ilqr_debug_cq_prof(buf)
a = np.array(list(buf), dtype=np.float64).reshape(2, 16, 2) / (B * (N + 1))
names = ("load", "walk", "sets", "vel", "terms", "rows", "rows2", "tables", "grad", "store")
print(f"B {B} cost quadratics {ms:.3f} ms per launch; cycles per knot, work/wait for warp A | warp B")
for p, n in enumerate(names):
    print(f"  {n:7s} {a[0, p, 0]:7.0f}/{a[0, p, 1]:6.0f} | {a[1, p, 0]:7.0f}/{a[1, p, 1]:6.0f}")
print(f"  total   {a[0].sum():7.0f}        | {a[1].sum():7.0f}")
