"""Exploratory timings on the GPU box (not the bench): single-instance latency, batch throughput, stage split."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_ilqr_mujoco_b200 import Config, gpu
from mpc_ilqr_mujoco_b200.references import ReferenceSet, perturbed_states, standing_state

w = Config().build_weights()
d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
s1 = gpu.H1IlqrBatch(w, N=25, batch=1)
print("fp64 peak TFLOP/s", s1.measure_fp64_peak(), "DMMA", s1.measure_fp64_mma_peak())
refs = ReferenceSet(d["walking_q"], d["walking_v"], d["walking_contact"], s1.reference_kinematics)
bias = s1.bias_forces(standing_state()[None])[0]
ug = np.zeros(19); ug[:18] = bias[7:25]
POLICY = int(os.environ.get("H1_POLICY", "0"))
for B in [int(a) for a in (sys.argv[1:] or ["1", "1024"])]:
    s = gpu.H1IlqrBatch(w, N=25, batch=B)
    s.set_kernel_policy(POLICY)
    s.set_reference_window(*refs.window(0, 25), shared=True)
    x0 = perturbed_states(standing_state(), B, seed=0)
    s.upload_inputs(x0, ug)
    for _ in range(2): s.run_resident_steps(1, True)
    ms = s.run_resident_steps(3, True) / 3
    ct, at = s.solve_trace()
    s.enable_stage_timing(True)
    s.mpc_reset(); s.initialize(x0, None, ug); c, it, st = s.solve(x0)
    tm = s.stage_times()
    s.enable_stage_timing(False)
    print(json.dumps({"policy": POLICY, "B": B, "ms_per_step": ms, "solves_per_s": B / ms * 1e3, "iters_mean": float(it.mean()), "iters_max": int(it.max()),
                      "stage_ms": {k: round(v, 3) for k, v in tm.items()}}))
    t0 = time.perf_counter(); s.mpc_reset(); ua, cc = s.mpc_step(x0, ug); t1 = time.perf_counter()
    print("  e2e mpc_step wall ms", (t1 - t0) * 1e3)
    s.close()
