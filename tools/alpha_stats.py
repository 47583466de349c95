"""Distribution of accepted line-search candidates over a batch (argv[1] instances) — sizing input for the
two-phase line search. Prints the histogram of first/second-attempt alpha indices (-1 = no candidate improved)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_ilqr_mujoco_b200 import Config, gpu
from mpc_ilqr_mujoco_b200 import workloads as wl
from mpc_ilqr_mujoco_b200.references import standing_state

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
s = gpu.H1IlqrBatch(Config().build_weights(), N=25, batch=B)
s.set_kernel_policy(int(os.environ.get("H1_POLICY", "0")))
win, x0, _ = wl.walking_instances(np.arange(B), s.reference_kinematics)
s.set_reference_window(*win, shared=False)
ug = np.zeros(19); ug[:18] = s.bias_forces(standing_state()[None])[0][7:25]
s.upload_inputs(x0, ug)
for _ in range(2):
    ms = s.run_resident_steps(1, True)
ct, at = s.solve_trace()
for k, name in ((0, "first attempt"), (1, "second attempt")):
    v = at[:, :, k].ravel(); v = v[v != -2]
    print(name, {int(a): int((v == a).sum()) for a in np.unique(v)})
print("B", B, "ms/step", ms, "solves/s", B / ms * 1e3, "launches", s.stage_times()["launches"])
s.enable_stage_timing(True)
s.mpc_reset(); s.initialize(x0, None, ug)
s.solve(x0)
print({k: round(v, 2) if isinstance(v, float) else v for k, v in s.stage_times().items()})
