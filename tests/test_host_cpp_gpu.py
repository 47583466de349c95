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


def _workdir(tmp_path, steps, tag="standing"):
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    (tmp_path / "data").mkdir()
    (tmp_path / "results").mkdir()
    from test_host import _unpack_models
    _unpack_models(tmp_path)     # robots/h1_description/{mjcf,urdf}: parsed at run time by RobotUtils::loadModel / iLQR
    np.savetxt(tmp_path / "data" / "q.csv", d[f"{tag}_q"], delimiter=",", fmt="%.17g")
    np.savetxt(tmp_path / "data" / "v.csv", d[f"{tag}_v"], delimiter=",", fmt="%.17g")
    with open(tmp_path / "data" / "c.csv", "w") as f:
        f.write("left_foot,right_foot\n")
        for r in d[f"{tag}_contact"]:
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
    assert "Parsed MJCF model" in out.stdout and "Parsed URDF cost model" in out.stdout    # run-time model loading
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


def _cxx_default(v):
    """A double as `std::ostream << v` prints it with default flags (precision 6, %g)."""
    return "%g" % v


def test_step_csv_and_optimal_trajectory_logs(tmp_path, oracle):
    """the wide step CSV (MPC::enableCSVLogging / logCurrentStep / finalizeCSVLog, mpc.cpp:181-268) and
    q_optimal.csv / u_optimal.csv (mpc.cpp:270-355) in the reference's byte format: header strings, one row per step
    written AFTER t_idx_ was advanced, default ostream number formatting (6 significant digits, no fixed notation),
    '\n' line ends — the files simulate.py:64-68 / plotter.py:33-46 parse. Values against the oracle's closed loop."""
    exe = os.path.join(HOST, "bin", "humanoid_mpc_demo")
    steps = 3
    _workdir(tmp_path, steps)
    out = subprocess.run([exe, "config.yaml", str(steps), "results/mpc_log.csv"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    raw = (tmp_path / "results" / "mpc_log.csv").read_bytes()
    assert b"\r" not in raw and raw.endswith(b"\n")
    lines = raw.decode().split("\n")[:-1]
    header = "time_index,time_sec,solve_cost,solve_time_ms" + "".join(f",x_{i}" for i in range(51)) + "".join(f",u_{i}" for i in range(19)) \
        + "".join(f",x_ref_{i}" for i in range(51)) + "".join(f",u_ref_{i}" for i in range(19))
    assert lines[0] == header and len(lines) == steps + 1
    so, w, win = make_oracle("standing")
    refs = reference_set("standing")
    x = standing_state(); ug = grav_comp_guess(x)
    for k in range(steps):
        tok = lines[1 + k].split(",")
        assert len(tok) == 4 + 51 + 19 + 51 + 19
        assert all(_cxx_default(float(t)) == t for t in tok), "not default ostream formatting"
        so.set_reference_window(*refs.window(k, 25))
        u, c = so.mpc_step(x, ug)
        assert tok[0] == str(k + 1) and tok[1] == _cxx_default((k + 1) * 0.02)      # logged after t_idx_++ (mpc.cpp:113-119)
        assert abs(float(tok[2]) - c) <= 2e-5 * abs(c)
        assert np.allclose([float(t) for t in tok[4:55]], x, rtol=2e-5, atol=2e-6)
        assert np.allclose([float(t) for t in tok[55:74]], u, rtol=2e-5, atol=2e-5)
        assert np.allclose([float(t) for t in tok[74:125]], refs.window(k, 25)[0][0], rtol=2e-5, atol=1e-12)
        x = oracle.dyn_step(x, u)[0]
    ql = (tmp_path / "results" / "q_optimal.csv").read_text().split("\n")[:-1]
    ul = (tmp_path / "results" / "u_optimal.csv").read_text().split("\n")[:-1]
    assert ql[0] == "step,time_sec" + "".join(f",q_{i}" for i in range(26)) and len(ql) == steps + 1
    assert ul[0] == "step,time_sec" + "".join(f",u_{i}" for i in range(19)) and len(ul) == steps + 1
    for row in ql[1:] + ul[1:]:
        assert all(_cxx_default(float(t)) == t for t in row.split(","))
    assert ql[1].split(",")[:2] == ["1", "0.02"] and ql[1].split(",")[4] == "1.0432"   # first knot of the optimised trajectory = x0


def test_robot_utils_remaining_public_surface(tmp_path, oracle):
    """The RobotUtils entry points main does not call (robot_utils.hpp:51-57, 74-81, 95-107), each one C-ABI call on the
    GPU: linearizeDynamicsFD, constraintCost / Gradients / Hessians, stageCost / terminalCost, getEEVelReference,
    getCoMVelReference with real targets, jointId, resetToReference, scaleRobotMass — against the CPU oracle."""
    import json
    from mpc_ilqr_mujoco_b200 import Config
    exe = os.path.join(HOST, "bin", "host_api_check")
    assert os.path.exists(exe), "host_api_check not built (make -C mpc-ilqr-mujoco_b200/host)"
    _workdir(tmp_path, 1, tag="walking")
    out = subprocess.run([exe, "config.yaml"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    x, u, T = np.array(r["x"]), np.array(r["u"]), r["t"]
    refs = reference_set("walking")
    assert np.abs(x[26:] - refs.x_ref_full[T][26:]).max() == 0.0          # resetToReference(T) + getState
    m = oracle.dynamics_model()
    w = Config().build_weights()
    # forward differences of f_D, eps 1e-5 (FD-vs-FD: see test_linearize_fd_parity for the bound)
    Ao, Bo = oracle.dyn_linearize(x, u, 1e-5)
    A = np.array(r["A"]).reshape(51, 51).T; B = np.array(r["B"]).reshape(19, 51).T
    sc = max(np.abs(Ao).max(), np.abs(Bo).max())
    assert np.abs(A - Ao).max() / sc < 1e-9 and np.abs(B - Bo).max() / sc < 1e-9
    # limit penalties: independent restatement of robot_utils.cpp:615-778
    def pen(v, lo, hi, wt):
        mg = 0.1 * (hi - lo); lo_s, hi_s = lo + mg, hi - mg
        if v > hi_s: return wt * (v - hi_s) ** 2, 2 * wt * (v - hi_s), 2 * wt
        if v < lo_s: return wt * (lo_s - v) ** 2, -2 * wt * (lo_s - v), 2 * wt
        return 0.0, 0.0, 0.0
    cr, jr = np.array(m.ctrl_range), np.array(m.jnt_range)
    pu = [pen(u[i], cr[i, 0], cr[i, 1], w.w_control_limits) for i in range(19)]
    pq = [pen(x[7 + i], jr[i, 0], jr[i, 1], w.w_joint_limits) for i in range(19)]
    cc = sum(p[0] for p in pu) + sum(p[0] for p in pq)
    assert cc > 0 and abs(r["constraint_cost"] - cc) <= 1e-12 * cc
    gx = np.zeros(51); gx[7:26] = [p[1] for p in pq]; hx = np.zeros(51); hx[7:26] = [p[2] for p in pq]
    assert np.abs(np.array(r["grad_x"]) - gx).max() <= 1e-12 * np.abs(gx).max()
    assert np.abs(np.array(r["grad_u"]) - [p[1] for p in pu]).max() <= 1e-12 * max(abs(p[1]) for p in pu)
    assert (np.array(r["hess_xx_diag"]) == hx).all() and (np.array(r["hess_uu_diag"]) == [p[2] for p in pu]).all()
    # stage / terminal cost (robot_utils.cpp:162-252)
    Q, R, Qf = np.array(w.Qdiag), np.array(w.Rdiag), np.array(w.Qfdiag)
    com = oracle.dyn_com(x)
    def stage(row):
        e = x - refs.x_ref_full[row]
        return 0.5 * e @ (Q * e) + 0.5 * u @ (R * u) + 0.5 * w.w_com * np.sum((com - refs.com_ref_full[row]) ** 2) + cc
    assert abs(r["stage_cost"] - stage(T)) <= 1e-11 * stage(T)
    assert abs(r["stage_cost_far"] - stage(refs.T - 1)) <= 1e-11 * stage(refs.T - 1)      # past the table: last row
    e = x - refs.x_ref_full[-1]
    term = 0.5 * e @ (Qf * e) + 0.5 * w.w_com * np.sum((com - refs.com_ref_full[-1]) ** 2) + sum(p[0] for p in pq)
    assert abs(r["terminal_cost"] - term) <= 1e-11 * term
    # whole symmetric matrices through setCostWeights (off-diagonal entries set in host_api_check.cpp), and back to diagonal
    eu = u
    extra = e0 = x - refs.x_ref_full[T]
    full_stage = stage(T) + (7.0 * e0[0] * e0[1] + -3.0 * e0[30] * e0[8]) + 0.0004 * eu[2] * eu[5]
    assert abs(r["stage_cost_fullq"] - full_stage) <= 1e-11 * abs(full_stage)
    full_term = term + 11.0 * e[10] * e[40]
    assert abs(r["terminal_cost_fullq"] - full_term) <= 1e-11 * abs(full_term)
    assert r["stage_cost_back"] == r["stage_cost"]
    # per-row velocity targets (robot_utils.cpp:388-412)
    xr = refs.x_ref_full[T]
    assert np.abs(np.array(r["com_vel"]) - oracle.dyn_com_vel(xr)).max() < 1e-12
    def flow(xx, eps):
        y = xx.copy(); y[0:3] += eps * xx[26:29]
        wv = xx[29:32]; ang = np.linalg.norm(wv) * eps
        b = np.array([np.cos(ang / 2), *(np.sin(ang / 2) * wv / max(np.linalg.norm(wv), 1e-300))]); a = xx[3:7]
        y[3:7] = [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                  a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]
        y[7:26] += eps * xx[32:]
        return y
    for ee, body in ((0, 5), (1, 10)):
        fd = (oracle.dyn_body_pos(flow(xr, 1e-6), body) - oracle.dyn_body_pos(flow(xr, -1e-6), body)) / 2e-6
        assert np.abs(np.array(r[f"ee_vel_{ee}"]) - fd).max() < 1e-8
    # plant step, and the plant after scaleRobotMass(1.5) (body masses only, robot_utils.cpp:835-842)
    assert np.abs(np.array(r["x_next"]) - oracle.dyn_step(x, u)[0]).max() < 1e-11
    for b in range(20):
        m.mass[b] *= 1.5
    m.total_mass *= 1.5
    assert np.abs(np.array(r["x_next_heavy"]) - oracle.dyn_step(x, u, model=m)[0]).max() < 1e-11
    assert np.abs(np.array(r["x_next_heavy"]) - np.array(r["x_next"])).max() > 1e-4
    assert (r["joint_id_torso"], r["joint_id_left_knee"], r["joint_id_unknown"]) == (11, 4, -1)
