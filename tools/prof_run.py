"""One cold MPC step at batch B (argv[1]) — the short command ncu wraps for kernel captures (tools/ncu_summary.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_ilqr_mujoco_b200 import Config, gpu
from mpc_ilqr_mujoco_b200.references import ReferenceSet, perturbed_states, standing_state

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = Config().build_weights()
d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
s = gpu.H1IlqrBatch(w, N=25, batch=B)
s.set_kernel_policy(int(os.environ.get("H1_POLICY", "0")))
refs = ReferenceSet(d["walking_q"], d["walking_v"], d["walking_contact"], s.reference_kinematics)
ug = np.zeros(19); ug[:18] = s.bias_forces(standing_state()[None])[0][7:25]
if os.environ.get("H1_PROF_WORKLOAD", "standing") == "bench":   # the bench.py workload (per-instance walking windows)
    from mpc_ilqr_mujoco_b200 import workloads as wl
    win, x0, _ = wl.walking_instances(np.arange(B), s.reference_kinematics)
    s.set_reference_window(*win, shared=False)
else:
    s.set_reference_window(*refs.window(0, 25), shared=True)
    x0 = perturbed_states(standing_state(), B, seed=0)
s.upload_inputs(x0, ug)
ms = s.run_resident_steps(1, True)
print("B", B, "ms", ms)
