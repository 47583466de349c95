"""Per-phase cycles of k_backward from a -DRIC_PROF build of the library (GPU box), averaged over ALL blocks:
for warp 0 (a contraction warp) and warp 7 (the sequential warp), work before / wait at each block-wide barrier, per knot.
Build first (here):  cd mpc-ilqr-mujoco_b200/csrc && nvcc $(Makefile flags) -DRIC_PROF -shared -o ../lib/libh1ilqr_prof.so h1ilqr_capi.cu model_tables.cpp -lcudart
usage: H1ILQR_LIB=$PWD/mpc-ilqr-mujoco_b200/lib/libh1ilqr_prof.so [H1_RIC_DENSE=1] python tools/ric_prof.py [B]"""
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
buf = (ctypes.c_ulonglong * 40)()
L.h1ilqr_debug_ric_prof(buf)                      # clear
reps = 3
ms = s.time_stage("backward", reps)
L.h1ilqr_debug_ric_prof(buf)
a = np.array(list(buf), dtype=np.float64).reshape(2, 20) / (B * N * reps)
phases = (("G5+copy (previous knot)", 0), ("G1b", 1), ("G3", 2), ("P3: G1a+G2 | LDL+Linv", 7), ("solves+G4 | copies", 4))
print(f"B {B} dense={os.environ.get('H1_RIC_DENSE', '0')} backward {ms:.3f} ms per launch; cycles per knot (work/wait)")
for w, name in ((0, "warp 0"), (1, "warp 7")):
    tot = sum(a[w][2 * p] + a[w][2 * p + 1] for _, p in phases)
    print(f"  {name}: " + "  ".join(f"{n} {a[w][2 * p]:.0f}/{a[w][2 * p + 1]:.0f}" for n, p in phases) + f"  total {tot:.0f}")
