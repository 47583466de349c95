#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the kernels of a small batched solve, a closed-loop step and the stand-alone
# queries (GPU box). usage: [SAN_B=24] [SAN_POLICIES=2,1] [SAN_QUICK=1] [SAN_TOOLS="memcheck racecheck"] tools/sanitize.sh TAG
# -> gpurun_out/TAG_sanitize_*.log. racecheck is slow on the warp-cooperative family (policy 1): B = 24 with the closed loop does not
# finish in 20 minutes; SAN_B=2 SAN_QUICK=1 (one cold solve + the queries) takes a few minutes.
TAG=${1:-san}; O=gpurun_out; mkdir -p $O
cat > /tmp/san_run.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from mpc_ilqr_mujoco_b200 import Config, gpu
from mpc_ilqr_mujoco_b200 import workloads as wl
from mpc_ilqr_mujoco_b200.references import standing_state
w = Config().build_weights()
B = int(os.environ.get('SAN_B', '24'))
quick = os.environ.get('SAN_QUICK', '0') == '1'
for policy in [int(v) for v in os.environ.get('SAN_POLICIES', '2,1').split(',')]:
    s = gpu.H1IlqrBatch(w, N=25, batch=B)
    s.set_kernel_policy(policy)
    win, x0, t0 = wl.walking_instances(np.arange(B) * 15, s.reference_kinematics)
    s.set_reference_window(*win, shared=False)
    ug = np.zeros(19); ug[:18] = s.bias_forces(standing_state()[None])[0][7:25]
    ua, c = s.mpc_step(x0, ug)
    out = {"rc": None}
    if not quick:
        ua, c = s.mpc_step(s.dynamics_step(x0, ua), ug)          # warm step
        refs = wl.reference_set("walking", s.reference_kinematics, s.reference_com_velocity)
        s.set_reference_table(refs)
        out = s.run_closed_loop(2, t_idx0=t0.astype(np.int32), x_start=x0, u_init=ug, graph=False)
    s.reference_ee_velocity(x0); s.limit_penalties(x0, ua); s.stage_cost(x0, ua, win[0][:, 0]); s.linearize_state(x0[0], ua[0], 1)
    s.get_cost_quadratics()
    print("policy", policy, "ok", np.isfinite(c).all(), out["rc"], flush=True)
    s.close()
PY
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_run.py > $O/${TAG}_sanitize_$tool.log 2>&1; echo "$tool rc $?"
  grep -E "policy|SUMMARY" $O/${TAG}_sanitize_$tool.log | tail -4
done
