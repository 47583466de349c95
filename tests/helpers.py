"""Shared test helpers: problem setup on the oracle side (CPU) used by both CPU and GPU tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mpc_ilqr_mujoco_b200 import Config  # noqa: E402
from mpc_ilqr_mujoco_b200.references import ReferenceSet, standing_state  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def oracle_kinematics(x):
    x = np.atleast_2d(x)
    com = np.array([po.dyn_com(r) for r in x])
    ee = np.array([[po.dyn_body_pos(r, 5), po.dyn_body_pos(r, 10)] for r in x])
    return com, ee


_REFS = {}


def reference_set(tag, kinematics=oracle_kinematics):
    key = (tag, kinematics)
    if key not in _REFS:
        d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
        _REFS[key] = ReferenceSet(d[f"{tag}_q"], d[f"{tag}_v"], d[f"{tag}_contact"], kinematics)
    return _REFS[key]


def grav_comp_guess(x0):
    """Cold-start guess of the reference (computeGravComp, robot_utils.cpp:844-866) incl. quirk Q15:
    torque i is read from qfrc_bias[7+i]; i = 18 reads one past the end, defined here as 0."""
    b = po.dyn_bias(x0)
    u = np.zeros(19)
    u[:18] = b[7:25]
    return u


def make_oracle(tag="standing", N=25, t0=0, batch=1, cfg=None, linearization=0):
    cfg = cfg or Config()
    w = cfg.build_weights()
    opt = po.default_options()
    opt.linearization = linearization  # 0 analytic (default), 1 forward differences (the reference's method)
    s = po.OracleSolver(w, N, batch=batch, options=opt)
    refs = reference_set(tag)
    win = refs.window(t0, N)
    s.set_reference_window(*win)
    return s, w, win
