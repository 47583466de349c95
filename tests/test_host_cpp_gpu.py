"""GPU test of the host-side C++ mirror of the reference's class API (mpc-ilqr-mujoco_b200/host): the demo binary
(RobotUtils / iLQR / MPC over the C ABI) must reproduce the oracle's closed-loop MPC costs, and — when it was
built in the build container — the UNMODIFIED reference main/humanoid_mpc.cpp linked against this library must run."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT, grav_comp_guess, make_oracle, reference_set, standing_state
from test_host import YAML

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "mpc-ilqr-mujoco_b200", "host")


def _workdir(tmp_path, steps):
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    (tmp_path / "data").mkdir()
    (tmp_path / "results").mkdir()
    np.savetxt(tmp_path / "data" / "q.csv", d["standing_q"], delimiter=",", fmt="%.17g")
    np.savetxt(tmp_path / "data" / "v.csv", d["standing_v"], delimiter=",", fmt="%.17g")
    with open(tmp_path / "data" / "c.csv", "w") as f:
        f.write("left_foot,right_foot\n")
        for r in d["standing_contact"]:
            f.write(f"{r[0]},{r[1]}\n")
    y = YAML.replace("data/q_ref2_mj.csv", "data/q.csv").replace("data/v_ref2.csv", "data/v.csv")
    y = y.replace("data/contact_walking.csv", "data/c.csv").replace("sim_steps: 100", f"sim_steps: {steps}")
    (tmp_path / "config.yaml").write_text(y)


def _oracle_costs(oracle, steps):
    so, w, win = make_oracle("standing")
    refs = reference_set("standing")
    x = standing_state()
    ug = grav_comp_guess(x)
    costs = []
    for k in range(steps):
        so.set_reference_window(*refs.window(k, 25))
        u, c = so.mpc_step(x, ug)
        costs.append(c)
        x = oracle.dyn_step(x, u)[0]
    return costs


def test_demo_binary_matches_oracle_closed_loop(tmp_path, oracle):
    exe = os.path.join(HOST, "bin", "humanoid_mpc_demo")
    assert os.path.exists(exe), "host demo not built (make -C mpc-ilqr-mujoco_b200/host)"
    steps = 4
    _workdir(tmp_path, steps)
    out = subprocess.run([exe, "config.yaml", str(steps)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    costs = [float(m) for m in re.findall(r"Cost: ([-+0-9.eE]+)", out.stdout)]
    ref = _oracle_costs(oracle, steps)
    assert len(costs) == steps
    assert np.allclose(costs, ref, rtol=2e-5), (costs, ref)   # stdout prints 6 significant digits


def test_unmodified_reference_main_runs(tmp_path, oracle):
    exe = os.path.join(HOST, "bin", "humanoid_mpc_reference_main")
    if not os.path.exists(exe):
        pytest.skip("compat binary is only built where /root/reference exists")
    steps = 3
    _workdir(tmp_path, steps)
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    costs = [float(m) for m in re.findall(r"Cost: ([-+0-9.eE]+)", out.stdout)]
    assert len(costs) == steps and np.isfinite(costs).all()
    assert abs(costs[0] - _oracle_costs(oracle, 1)[0]) <= 2e-5 * abs(costs[0])
    q = (tmp_path / "results" / "q_optimal.csv").read_text().splitlines()
    assert q[0].startswith("step,time_sec,q_0") and len(q) == steps + 1
