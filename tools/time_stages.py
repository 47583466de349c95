"""Per-stage device times on the bench workload (GPU box): B instances (argv[1], default 8192) are initialised, rolled
out, linearized and given gains exactly as the first iteration of a cold MPC step does; every stage is then timed alone
(CUDA events, h1ilqr_time_stage). usage: python tools/time_stages.py [B] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_ilqr_mujoco_b200 import Config, gpu  # noqa: E402
from mpc_ilqr_mujoco_b200 import workloads as wl  # noqa: E402
from mpc_ilqr_mujoco_b200.references import standing_state  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = Config().build_weights()
s = gpu.H1IlqrBatch(w, N=25, batch=B)
win, x0, _ = wl.walking_instances(np.arange(B), s.reference_kinematics)
s.set_reference_window(*win, shared=False)
ug = np.zeros(19); ug[:18] = s.bias_forces(standing_state()[None])[0][7:25]
s.initialize(x0, None, ug)
s.rollout_nominal(x0); s.linearize(); s.cost_quadratics(); s.backward_pass()
out = {}
for st in ("factor", "linearize", "cost_quadratics", "backward", "line_search"):
    s.time_stage(st, 1)
    out[st] = round(s.time_stage(st, reps), 4)
print("B", B, "ms per launch", out)
