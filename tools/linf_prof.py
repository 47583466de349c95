"""Cycles of a warp between the numbered steps of k_linearize_finish from a -DLINF_PROF build (GPU box), averaged over all knots.
Build first (here): cd mpc-ilqr-mujoco_b200/csrc && nvcc $(Makefile flags) -DLINF_PROF -shared -o ../lib/libh1ilqr_prof.so h1ilqr_capi.cu model_tables.cpp -lcudart
usage: H1ILQR_LIB=$PWD/mpc-ilqr-mujoco_b200/lib/libh1ilqr_prof.so python tools/linf_prof.py [B]"""
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
buf = (ctypes.c_ulonglong * 8)()
L.h1ilqr_debug_linf_prof(buf)
ms = s.time_stage("linearize", 1)
L.h1ilqr_debug_linf_prof(buf)
a = np.array(list(buf), dtype=np.float64) / (B * N)
names = ("issue copies + zero fill", "wait factor", "N = L^-1", "Mhat^-1", "wait tangents", "Adot = Mhat^-1 T", "integrator + stores")
print(f"B {B} linearize {ms:.3f} ms; k_linearize_finish cycles per knot (one warp): total {a[:7].sum():.0f}")
for n, v in zip(names, a):
    print(f"  {n:26s} {v:7.0f}")
