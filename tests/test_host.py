"""CPU tests of the host-side logic and of the C-ABI library surface (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, reference_set, standing_state
from mpc_ilqr_mujoco_b200 import Config, load_config_from_file
from mpc_ilqr_mujoco_b200.references import load_contact_csv, load_qv_csv, perturbed_states
from mpc_ilqr_mujoco_b200.sharding import shard_range

YAML = """
robot:
  name: h1
  model_path: "robots/h1_description/mjcf/scene.xml"
  urdf_path: "robots/h1_description/urdf/h1.urdf"
  ee_feet:
    right_feet_ee: "right_ankle_link"
reference_trajectory:
  q_ref: "data/q_ref2_mj.csv"
  v_ref: "data/v_ref2.csv"
  contact_schedule: "data/contact_walking.csv"
mpc:
  horizon: 25              # N
  dt: 0.02
  physics_dt: 0.02
  gravity: [0.0, 0.0, -1.0]
  sim_steps: 100
  contact_impratio: 100.0
  cost_weights:
    Q_position_x: 200.0
    Q_position_y: 50.0
    Q_position_z: 200.0
    Q_quat_w: 50.0
    Q_quat_xyz: [50.0, 50.0, 50.0]
    Q_joint_pos: 50.0
    Q_vel_x: 150.0
    Q_vel_y: 50.0
    Q_vel_z: 150.0
    Q_ang_vel: 75.0
    Q_joint_vel: 75.0
    R_control: 0.001
    Qf_multiplier: 2.0
    Qf_position_x: 5.0
    Qf_position_y: 2.0
    Qf_position_z: 5.0
    Qf_vel_z: 4.0
    W_com_pos: 100.0
    W_com_vel: 0.0
    W_foot: 400.0
    W_foot_vel: 400.0
    W_upright: 20.0
    w_balance: 30.0
  constraints:
    joint_limit_weight: 1500.0
    torque_limit_weight: 1500.0
paths:
  logs_dir: "logs"
logging:
  verbose: true
  save_trajectories: true
  results_path: "results"
"""


def test_config_yaml_and_cost_matrices(tmp_path):
    p = tmp_path / "config.yaml"
    p.write_text(YAML)
    cfg = load_config_from_file(str(p))
    assert cfg.mpc.horizon == 25 and cfg.mpc.gravity == [0.0, 0.0, -1.0] and cfg.mpc.costs.W_com == 100.0
    w = cfg.build_weights()
    Q, Qf, R = np.array(w.Qdiag), np.array(w.Qfdiag), np.array(w.Rdiag)
    assert list(Q[:7]) == [200, 50, 200, 50, 50, 50, 50] and (Q[7:26] == 50).all()
    assert list(Q[26:32]) == [150, 50, 150, 75, 75, 75] and (Q[32:] == 75).all()
    assert (R == 0.001).all()
    exp = 2 * Q; exp[0] *= 5; exp[1] *= 2; exp[2] *= 5; exp[28] *= 4   # config.cpp:108-117
    assert np.allclose(Qf, exp)
    assert (w.w_ee_pos, w.w_ee_vel, w.w_upright, w.w_balance) == (400.0, 400.0, 20.0, 30.0)
    d = Config().build_weights()   # defaults == shipped config.yaml
    assert np.allclose(np.array(d.Qfdiag), Qf) and d.w_joint_limits == 1500.0
    with pytest.raises(KeyError):
        bad = tmp_path / "bad.yaml"; bad.write_text(YAML.replace("horizon: 25", "horizonx: 25"))
        load_config_from_file(str(bad))


def test_csv_loaders_skip_rules(tmp_path):
    q = tmp_path / "q.csv"; v = tmp_path / "v.csv"; c = tmp_path / "c.csv"
    row_q = ",".join(["0.5"] * 26); row_v = ",".join(["0.25"] * 25)
    q.write_text(row_q + "\n" + ",".join(["1"] * 20) + "\n" + row_q)   # 2nd row: wrong column count -> skipped
    v.write_text(row_v + "\n" + row_v + "\n" + row_v + "\n")
    c.write_text("left_foot,right_foot\n1,1\n0,1\n\n1,0\n")
    Q, V = load_qv_csv(str(q), str(v))
    assert Q.shape == (2, 26) and V.shape == (2, 25)
    assert load_contact_csv(str(c)).tolist() == [[1, 1], [0, 1], [1, 0]]


def test_reference_windows_and_quirk_q6():
    refs = reference_set("walking")
    assert refs.T == 400 and refs.contact.shape == (400, 2)
    x_ref, u_ref, com_ref, ee_ref, stance, cv = refs.window(390, 25)
    assert (x_ref[10:] == refs.x_ref_full[399]).all() and (u_ref == 0).all()   # clamped at the last row
    # Q6: flags / foot targets use the horizon-LOCAL index: identical for every window start
    w0, w1 = refs.window(0, 25), refs.window(100, 25)
    assert (w0[4] == w1[4]).all() and (w0[3] == w1[3]).all() and not (w0[0] == w1[0]).all()
    assert stance[:22].tolist() == [[1, 1]] * 22 and stance[22:26].tolist() == [[0, 1]] * 4
    w2 = refs.window(100, 25, schedule_offset=True)
    assert (w2[4] == refs.contact[100:126]).all()
    st = reference_set("standing")
    assert np.abs(st.com_ref_full[0] - [0.0162336940, 0.000967503071, 1.00406373]).max() < 5e-9
    assert refs.is_stance(0, 10_000) == 1 and refs.is_stance(5, 0) == 1   # defaults outside the schedule


def test_perturbed_states_are_deterministic_and_bounded():
    a = perturbed_states(standing_state(), 5, seed=0)
    b = perturbed_states(standing_state(), 5, seed=0)
    assert (a == b).all() and not (a[0] == a[1]).all()
    assert np.abs(np.linalg.norm(a[:, 3:7], axis=1) - 1).max() < 1e-15
    assert np.abs(a[:, :3] - standing_state()[:3]).max() <= 0.02 and np.abs(a[:, 26:]).max() <= 0.1


def test_shard_range_partitions_everything():
    for total, world in ((65536, 8), (10, 4), (3, 8)):
        blocks = [shard_range(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1


def test_c_abi_library_exports_every_declared_symbol():
    """The shared library must load here (no GPU) and export exactly what include/h1ilqr.h declares."""
    from mpc_ilqr_mujoco_b200 import gpu
    L = gpu.lib()
    hdr = open(os.path.join(ROOT, "include", "h1ilqr.h")).read() + open(os.path.join(ROOT, "include", "h1_model.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(h1ilqr_\w+|h1_default_\w+)\s*\(", hdr))
    assert len(names) >= 35
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/ but not exported"
    assert names == set(gpu.EXPORTS)
    o = gpu.default_options()
    assert o.max_iterations == 10 and o.tolerance == 1e-4 and o.fd_eps == 1e-5 and list(o.alphas)[:3] == [1.0, 0.8, 0.6]
    assert abs(sum(gpu.default_dynamics_model().mass) - 51.649896) < 1e-9


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mpc_ilqr_mujoco_b200 import gpu
    with pytest.raises(gpu.H1IlqrError):
        gpu.H1IlqrBatch(Config().build_weights(), N=25, batch=1)


def test_pinocchio_order_reference_loader(tmp_path):
    """h1_walking_pin.csv (BASELINE config 2) is stored in Pinocchio order; its rows convert exactly to the MuJoCo-ordered
    q_ref2_mj.csv that the shipped config uses (get_contacts.py:18-41). Fixture: the first 64 rows of the file."""
    from mpc_ilqr_mujoco_b200.references import load_q_pin_csv, pinocchio_to_mujoco_q
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    pin = d["walking_pin_q_head"]
    assert pin.shape == (64, 26) and abs(np.linalg.norm(pin[0, 3:7]) - 1) < 1e-5 and pin[0, 6] > 0.99    # qw last
    mj = pinocchio_to_mujoco_q(pin)
    assert np.abs(mj - d["walking_q"][:64]).max() == 0.0
    assert (pinocchio_to_mujoco_q(pin[5]) == mj[5]).all()
    p = tmp_path / "pin.csv"
    np.savetxt(p, pin, delimiter=",", fmt="%.6f")
    with open(p, "a") as f:
        f.write("1,2,3\n")                                   # malformed row: skipped
    assert np.abs(load_q_pin_csv(str(p)) - mj).max() < 1e-12


def _unpack_models(tmp_path):
    """The reference's scene.xml / h1.xml / h1.urdf from data/h1_models.npz laid out like robots/h1_description/."""
    d = np.load(os.path.join(ROOT, "data", "h1_models.npz"))
    mj = tmp_path / "robots" / "h1_description" / "mjcf"; ur = tmp_path / "robots" / "h1_description" / "urdf"
    mj.mkdir(parents=True); ur.mkdir(parents=True)
    (mj / "scene.xml").write_bytes(d["mjcf_scene_xml"].tobytes()); (mj / "h1.xml").write_bytes(d["mjcf_h1_xml"].tobytes())
    (ur / "h1.urdf").write_bytes(d["urdf_h1_urdf"].tobytes())
    return str(mj / "scene.xml"), str(ur / "h1.urdf")


def test_runtime_model_loader_matches_generated_tables(tmp_path):
    """RobotUtils::loadModel / iLQR parse the MJCF (through scene.xml's <include>) and the URDF at run time
    (host/src/model_loader.cpp; reference: robot_utils.cpp:19-55, derivatives.cpp:26-39). The parsed H1Model tables equal
    the ones tools/gen_h1_model.py generated from the same files at build time; a non-H1 file is rejected."""
    import json
    import subprocess
    from mpc_ilqr_mujoco_b200 import gpu
    exe = os.path.join(ROOT, "mpc-ilqr-mujoco_b200", "host", "bin", "model_load_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "mpc-ilqr-mujoco_b200", "host")], stdout=subprocess.DEVNULL)
    scene, urdf = _unpack_models(tmp_path)
    out = subprocess.run([exe, scene, urdf], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout)
    assert r["joints"][0] == "left_hip_yaw_joint" and r["joints"][10] == "torso_joint" and len(r["joints"]) == 19
    for key, ref in (("dynamics", gpu.default_dynamics_model()), ("cost", gpu.default_cost_model())):
        m = r[key]
        assert m["parent"] == list(ref.parent) and m["axis"] == list(ref.axis) and m["has_rfix"] == list(ref.has_rfix)
        assert m["foot_body"] == list(ref.foot_body) == [5, 10]
        for name, val in (("pos", ref.pos), ("rfix", ref.rfix), ("mass", ref.mass), ("ipos", ref.ipos), ("inertia", ref.inertia),
                          ("armature", ref.armature), ("damping", ref.damping), ("jnt_range", ref.jnt_range),
                          ("ctrl_range", ref.ctrl_range), ("foot_pts", ref.foot_pts), ("gravity", ref.gravity)):
            a, b = np.array(m[name]), np.array(val).ravel()
            assert np.abs(a - b).max() <= 1e-15 * max(np.abs(b).max(), 1.0), (key, name)
        assert np.allclose(m["scalars"], [ref.timestep, ref.contact_kn, ref.contact_bn, ref.contact_bt, ref.contact_eps, ref.total_mass], rtol=1e-15)
    # a file that is not an H1-class model
    bad = tmp_path / "bad.xml"
    bad.write_text('<mujoco><worldbody><body name="a"><inertial pos="0 0 0" mass="1" diaginertia="1 1 1"/><freejoint/></body></worldbody></mujoco>')
    out = subprocess.run([exe, str(bad), urdf], capture_output=True, text=True)
    assert out.returncode == 1 and "expected 20 bodies" in out.stderr
