"""BASELINE config 4: H1 walking, horizon sweep N = 25 / 50 / 100 / 200, the 8 line-search candidates evaluated concurrently.
Prints one JSON line per horizon: single-instance ms per cold MPC step (latency kernels) and batched solves/s (1024 instances)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_ilqr_mujoco_b200 import Config, gpu
from mpc_ilqr_mujoco_b200.references import ReferenceSet, perturbed_states, standing_state

w = Config().build_weights()
d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
probe = gpu.H1IlqrBatch(w, N=25, batch=1)
refs = ReferenceSet(d["walking_q"], d["walking_v"], d["walking_contact"], probe.reference_kinematics)
ug = np.zeros(19); ug[:18] = probe.bias_forces(standing_state()[None])[0][7:25]
for N in (25, 50, 100, 200):
    out = {"horizon": N}
    for B in (1, 1024):
        s = gpu.H1IlqrBatch(w, N=N, batch=B)
        s.set_reference_window(*refs.window(0, N), shared=True)
        x0 = perturbed_states(refs.x_ref_full[0], B, seed=0)
        s.upload_inputs(x0, ug)
        for _ in range(2):
            s.run_resident_steps(1, True)
        ms = s.run_resident_steps(3, True) / 3
        ct, at = s.solve_trace()
        iters = float((at[:, :, 0] != -2).sum(axis=1).mean())
        if B == 1:
            out["single_instance_ms_per_mpc_step"] = ms; out["single_instance_iterations"] = iters
        else:
            out["batch"] = B; out["batched_solves_per_s"] = B / ms * 1e3; out["batched_mean_iterations"] = iters
        s.close()
    print(json.dumps(out), flush=True)
