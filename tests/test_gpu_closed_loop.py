"""Device-resident closed loop (main/humanoid_mpc.cpp:130-179, robot_utils.cpp:99-103, 422-443): window
extraction at the per-instance time index, warm start, solve, first control, plant step and time-index advance all on
the device (h1ilqr_set_reference_table + h1ilqr_run_closed_loop), checked against the oracle's closed loop driven from the
host step by step, and the CUDA-graph replay of a step against the plain launch sequence."""
import numpy as np
import pytest

from helpers import grav_comp_guess, make_oracle, oracle_kinematics, po, reference_set, rel_err, standing_state
from mpc_ilqr_mujoco_b200 import Config
from mpc_ilqr_mujoco_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from mpc_ilqr_mujoco_b200 import gpu as g
    g.lib()
    return g


def _oracle_loop(tag, t0, x_start, steps, ug):
    so, w, _ = make_oracle(tag)
    refs = reference_set(tag)
    x = x_start.copy()
    costs, us, its = [], [], []
    for k in range(steps):
        so.set_reference_window(*refs.window(t0 + k, 25))
        u, c = so.mpc_step(x, ug)
        costs.append(c); us.append(u.copy()); its.append(so.iters())
        x = po.dyn_step(x, u)[0]
    return x, np.array(costs), np.array(us), np.array(its)


def test_closed_loop_standing_15_steps(gpu, oracle):
    """BASELINE config 1 (standing, 15 MPC steps) for a small batch: instance 0 is the unperturbed standing pose (what
    main runs), the others start from perturbed poses; every instance against its own oracle loop."""
    w = Config().build_weights()
    B, steps = 4, 15
    sg = gpu.H1IlqrBatch(w, N=25, batch=B)
    refs = wl.reference_set("standing", sg.reference_kinematics, sg.reference_com_velocity)
    sg.set_reference_table(refs)
    jr = np.array(gpu.default_dynamics_model().jnt_range)
    x0 = np.vstack([standing_state()] + [wl.perturb(standing_state(), i, jnt_range=jr) for i in range(1, B)])
    ug = grav_comp_guess(standing_state())
    out = sg.run_closed_loop(steps, t_idx0=0, x_start=x0, u_init=ug, graph=False)
    assert out["rc"] == 0 and np.isfinite(out["x_final"]).all()
    for i in range(B):
        xf, costs, us, its = _oracle_loop("standing", 0, x0[i], steps, ug)
        assert (out["iters"][:, i] == its).all(), (i, out["iters"][:, i], its)
        assert rel_err(out["cost"][:, i], costs) < 1e-6, i
        assert np.abs(out["u"][:, i] - us).max() <= 1e-6 * max(np.abs(us).max(), 1.0), i
        assert np.abs(out["x_final"][i] - xf).max() <= 1e-6 * max(np.abs(xf).max(), 1.0), i


def test_closed_loop_walking_with_time_indices_and_graph(gpu, oracle):
    """Walking reference (config 2), 12 steps, per-instance time indices incl. one that crosses the end of the table
    (window clamp, robot_utils.cpp:430-441); the CUDA-graph replay must give bit-identical results to the plain launch
    sequence, and both must match the oracle loop for the benign starts (the mid-table starts are ill-conditioned, see
    test_config2_closed_loop: decisions and first steps are compared there with per-step re-synchronisation)."""
    w = Config().build_weights()
    steps = 12
    t0 = np.array([0, 0, 370, 120], dtype=np.int32)
    B = len(t0)
    refs_o = reference_set("walking")
    x0 = np.vstack([standing_state(), wl.perturb(standing_state(), 5), refs_o.x_ref_full[370], refs_o.x_ref_full[120]])
    ug = grav_comp_guess(standing_state())
    outs = []
    for graph in (False, True):
        sg = gpu.H1IlqrBatch(w, N=25, batch=B)
        refs = wl.reference_set("walking", sg.reference_kinematics)
        sg.set_reference_table(refs)
        outs.append(sg.run_closed_loop(steps, t_idx0=t0, x_start=x0, u_init=ug, graph=graph))
        # continuing the loop (no x_start / t_idx0) carries on from the device state
        more = sg.run_closed_loop(3, graph=graph)
        outs[-1]["more"] = more
        sg.close()
    a, b = outs
    assert (a["x_final"] == b["x_final"]).all() and (a["cost"] == b["cost"]).all() and (a["iters"] == b["iters"]).all()
    assert (a["more"]["x_final"] == b["more"]["x_final"]).all()
    for i in (0, 1):
        xf, costs, us, its = _oracle_loop("walking", int(t0[i]), x0[i], steps, ug)
        assert (a["iters"][:, i] == its).all(), (i, a["iters"][:, i], its)
        assert rel_err(a["cost"][:, i], costs) < 1e-6, i
        assert np.abs(a["x_final"][i] - xf).max() <= 1e-6 * max(np.abs(xf).max(), 1.0), i
    # first step of the hard starts: identical inputs on both sides, incl. the clamped window at t_idx 370
    for i in (2, 3):
        xf, costs, us, its = _oracle_loop("walking", int(t0[i]), x0[i], 2, ug)
        assert a["iters"][0, i] == its[0] and abs(a["cost"][0, i] - costs[0]) <= 1e-6 * abs(costs[0]), i
        assert np.abs(a["u"][0, i] - us[0]).max() <= 1e-6 * max(np.abs(us[0]).max(), 1.0), i


def test_resident_steps_graph_equals_stream(gpu):
    """h1ilqr_run_resident_steps: replaying the captured step gives the same solution as enqueueing its launches."""
    w = Config().build_weights()
    B = 64
    res = []
    for graph in (False, True):
        sg = gpu.H1IlqrBatch(w, N=25, batch=B)
        win, x0, _ = wl.walking_instances(np.arange(B), sg.reference_kinematics)
        sg.set_reference_window(*win, shared=False)
        sg.upload_inputs(x0, grav_comp_guess(standing_state()))
        ms = sg.run_resident_steps(2, True, graph=graph)
        assert ms > 0
        res.append(sg.get_trajectory())
        sg.close()
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all()
