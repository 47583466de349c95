"""Per-item cycles of k_linearize_tangents from a -DLINT_PROF build (GPU box): measured cost of every direction item, the busy
share of the eight warps of a CTA (sum of item times / 8 x CTA makespan) — the numbers DynModel::tan_order is tuned with.
Build first (here): cd mpc-ilqr-mujoco_b200/csrc && nvcc $(Makefile flags) -DLINT_PROF -shared -o ../lib/libh1ilqr_prof.so h1ilqr_capi.cu model_tables.cpp -lcudart
usage: H1ILQR_LIB=$PWD/mpc-ilqr-mujoco_b200/lib/libh1ilqr_prof.so [H1_LINT_SPLIT=1] python tools/lint_prof.py [B]"""
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
buf = (ctypes.c_ulonglong * 131)()
L.h1ilqr_debug_lint_prof(buf)
ms = s.time_stage("linearize", 1)
L.h1ilqr_debug_lint_prof(buf)
a = np.array(list(buf), dtype=np.float64)
groups = B * N / 32
item = a[:128] / groups
busy, span = a[128] / groups, a[129] / groups
print(f"B {B} split={os.environ.get('H1_LINT_SPLIT', '0')} linearize {ms:.3f} ms; per knot group: sum of item cycles {item.sum():.0f}, warp-busy {busy:.0f}, "
      f"CTA makespan(s) {span:.0f}, busy share {busy / (8 * span):.3f}")
for cls in range(3):
    print(f"  class {cls}: " + " ".join(f"{int(item[(cls << 5) | i])}" for i in range(22) if item[(cls << 5) | i] > 0))
