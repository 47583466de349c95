"""Host-side configuration: mirrors the reference's Config / loadConfigFromFile / buildCostMatrices.

Reference: include/common/config.hpp:9-51, src/common/config.cpp:4-122, config.yaml.
The YAML schema is unchanged; `build_weights` produces the DIAGONALS of Q, R, Qf exactly as
Config::buildCostMatrices fills them (config.cpp:66-117) plus the scalar task weights that
main/humanoid_mpc.cpp:104-111 hands to RobotUtils::set*Weight.
"""
from dataclasses import dataclass, field
from typing import List

import yaml

from .ctypes_defs import NQ, NU, NX, H1Weights


@dataclass
class CostWeights:
    Q_position_x: float = 200.0
    Q_position_y: float = 50.0
    Q_position_z: float = 200.0
    Q_quat_w: float = 50.0
    Q_quat_xyz: List[float] = field(default_factory=lambda: [50.0, 50.0, 50.0])
    Q_joint_pos: float = 50.0
    Q_vel_x: float = 150.0
    Q_vel_y: float = 50.0
    Q_vel_z: float = 150.0
    Q_ang_vel: float = 75.0
    Q_joint_vel: float = 75.0
    R_control: float = 0.001
    Qf_multiplier: float = 2.0
    Qf_position_x: float = 5.0
    Qf_position_y: float = 2.0
    Qf_position_z: float = 5.0
    Qf_vel_z: float = 4.0
    W_com: float = 100.0
    W_com_vel: float = 0.0
    W_foot: float = 400.0
    W_foot_vel: float = 400.0
    W_upright: float = 20.0
    w_balance: float = 30.0


@dataclass
class MpcParams:
    horizon: int = 25
    dt: float = 0.02
    physics_dt: float = 0.02
    gravity: List[float] = field(default_factory=lambda: [0.0, 0.0, -1.0])
    sim_steps: int = 100
    contact_impratio: float = 100.0
    costs: CostWeights = field(default_factory=CostWeights)
    joint_limit_weight: float = 1500.0
    torque_limit_weight: float = 1500.0


@dataclass
class Config:
    """Defaults equal the shipped config.yaml of the reference."""
    model_path: str = "robots/h1_description/mjcf/scene.xml"
    urdf_path: str = "robots/h1_description/urdf/h1.urdf"
    q_ref_path: str = "data/q_ref2_mj.csv"
    v_ref_path: str = "data/v_ref2.csv"
    contact_schedule_path: str = "data/contact_walking.csv"
    results_path: str = "results"
    verbose: bool = True
    save_trajectories: bool = True
    mpc: MpcParams = field(default_factory=MpcParams)

    def build_weights(self) -> H1Weights:
        """Config::buildCostMatrices (config.cpp:66-122) + main:104-111, as the C-ABI weight struct."""
        c = self.mpc.costs
        w = H1Weights()
        Q = [1.0] * NX
        Q[0], Q[1], Q[2], Q[3] = c.Q_position_x, c.Q_position_y, c.Q_position_z, c.Q_quat_w
        Q[4], Q[5], Q[6] = c.Q_quat_xyz
        for i in range(7, NQ):
            Q[i] = c.Q_joint_pos
        Q[NQ + 0], Q[NQ + 1], Q[NQ + 2] = c.Q_vel_x, c.Q_vel_y, c.Q_vel_z
        Q[NQ + 3] = Q[NQ + 4] = Q[NQ + 5] = c.Q_ang_vel
        for i in range(NQ + 6, NX):
            Q[i] = c.Q_joint_vel
        Qf = [q * c.Qf_multiplier for q in Q]
        Qf[0] *= c.Qf_position_x
        Qf[1] *= c.Qf_position_y
        Qf[2] *= c.Qf_position_z
        Qf[NQ + 2] *= c.Qf_vel_z
        for i in range(NX):
            w.Qdiag[i], w.Qfdiag[i] = Q[i], Qf[i]
        for i in range(NU):
            w.Rdiag[i] = 1.0 * c.R_control
        w.w_com, w.w_com_vel, w.w_ee_pos, w.w_ee_vel = c.W_com, c.W_com_vel, c.W_foot, c.W_foot_vel
        w.w_upright, w.w_balance = c.W_upright, c.w_balance
        w.w_joint_limits, w.w_control_limits = self.mpc.joint_limit_weight, self.mpc.torque_limit_weight
        return w


def load_config_from_file(path: str) -> Config:
    """loadConfigFromFile (config.cpp:4-64). Keys present in the YAML but never read by the reference
    (robot.name, robot.ee_feet, paths.*) are ignored here too. Missing keys raise KeyError (the
    reference exits on a YAML exception)."""
    with open(path) as f:
        y = yaml.safe_load(f)
    cfg = Config()
    cfg.model_path = str(y["robot"]["model_path"])
    cfg.urdf_path = str(y["robot"]["urdf_path"])
    rt = y["reference_trajectory"]
    cfg.q_ref_path, cfg.v_ref_path = str(rt["q_ref"]), str(rt["v_ref"])
    cfg.contact_schedule_path = str(rt["contact_schedule"])
    lg = y["logging"]
    cfg.results_path, cfg.verbose, cfg.save_trajectories = str(lg["results_path"]), bool(lg["verbose"]), bool(lg["save_trajectories"])
    m = y["mpc"]
    p = cfg.mpc
    p.horizon, p.dt, p.physics_dt = int(m["horizon"]), float(m["dt"]), float(m["physics_dt"])
    p.gravity = [float(g) for g in m["gravity"]]
    p.sim_steps, p.contact_impratio = int(m["sim_steps"]), float(m["contact_impratio"])
    cw = m["cost_weights"]
    c = p.costs
    for k in ("Q_position_x", "Q_position_y", "Q_position_z", "Q_quat_w", "Q_joint_pos", "Q_vel_x", "Q_vel_y",
              "Q_vel_z", "Q_ang_vel", "Q_joint_vel", "R_control", "Qf_multiplier", "Qf_position_x", "Qf_position_y",
              "Qf_position_z", "Qf_vel_z", "W_com_vel", "W_foot", "W_foot_vel", "W_upright", "w_balance"):
        setattr(c, k, float(cw[k]))
    c.Q_quat_xyz = [float(v) for v in cw["Q_quat_xyz"]]
    c.W_com = float(cw["W_com_pos"])  # YAML name differs from the struct field (config.cpp:47)
    p.joint_limit_weight = float(m["constraints"]["joint_limit_weight"])
    p.torque_limit_weight = float(m["constraints"]["torque_limit_weight"])
    return cfg


def dump_config_yaml(cfg: Config) -> str:
    """The reference's config.yaml schema (config.cpp:4-64) for `cfg` — what loadConfigFromFile / load_config_from_file
    read back."""
    c, m = cfg.mpc.costs, cfg.mpc
    lines = [
        "robot:", "  name: h1", f'  model_path: "{cfg.model_path}"', f'  urdf_path: "{cfg.urdf_path}"',
        "reference_trajectory:", f'  q_ref: "{cfg.q_ref_path}"', f'  v_ref: "{cfg.v_ref_path}"',
        f'  contact_schedule: "{cfg.contact_schedule_path}"',
        "mpc:", f"  horizon: {m.horizon}", f"  dt: {m.dt}", f"  physics_dt: {m.physics_dt}",
        f"  gravity: [{m.gravity[0]}, {m.gravity[1]}, {m.gravity[2]}]", f"  sim_steps: {m.sim_steps}",
        f"  contact_impratio: {m.contact_impratio}", "  cost_weights:"]
    for k in ("Q_position_x", "Q_position_y", "Q_position_z", "Q_quat_w"):
        lines.append(f"    {k}: {getattr(c, k)}")
    lines.append(f"    Q_quat_xyz: [{c.Q_quat_xyz[0]}, {c.Q_quat_xyz[1]}, {c.Q_quat_xyz[2]}]")
    for k in ("Q_joint_pos", "Q_vel_x", "Q_vel_y", "Q_vel_z", "Q_ang_vel", "Q_joint_vel", "R_control", "Qf_multiplier",
              "Qf_position_x", "Qf_position_y", "Qf_position_z", "Qf_vel_z"):
        lines.append(f"    {k}: {getattr(c, k)}")
    lines += [f"    W_com_pos: {c.W_com}", f"    W_com_vel: {c.W_com_vel}", f"    W_foot: {c.W_foot}", f"    W_foot_vel: {c.W_foot_vel}",
              f"    W_upright: {c.W_upright}", f"    w_balance: {c.w_balance}", "  constraints:",
              f"    joint_limit_weight: {m.joint_limit_weight}", f"    torque_limit_weight: {m.torque_limit_weight}",
              "logging:", f'  results_path: "{cfg.results_path}"', f"  verbose: {'true' if cfg.verbose else 'false'}",
              f"  save_trajectories: {'true' if cfg.save_trajectories else 'false'}", ""]
    return "\n".join(lines)
