"""Synthetic workloads of BASELINE.json's configurations, shared by bench.py and the parity tests
so that what is measured is exactly what is checked against the oracle.

  config 3 : `standing_instances(ids)`  — standing reference (one shared window), x0_i = standing pose + perturbation
  config 5 : `walking_instances(ids)`   — instance i tracks the walking reference from window row t0_i = i mod 374
                                          (= T - (N + 1)), x0_i = x_ref[t0_i] + perturbation, per-instance windows

The perturbation is counter-based (Philox, key = seed, counter = GLOBAL instance id) so that any subset of instances
can be generated on any rank: base xyz U(+-0.02 m), base orientation exp(U(+-0.05 rad)^3), joints U(+-0.05 rad)
(config 3 additionally clips them to the inner 80 % of the joint range), all 25 velocities U(+-0.1).
"""
import os

import numpy as np

from .ctypes_defs import NQ, NV
from .references import ReferenceSet, standing_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_HORIZON = 25


def perturb(x_nom, gid, seed=0, jnt_range=None):
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[int(gid), 0, 0, 0]))
    x = np.array(x_nom, dtype=np.float64)
    x[0:3] += rng.uniform(-0.02, 0.02, 3)
    rv = rng.uniform(-0.05, 0.05, 3)
    ang = np.linalg.norm(rv)
    dq = np.array([np.cos(ang / 2), *(np.sin(ang / 2) / ang * rv)]) if ang > 0 else np.array([1.0, 0.0, 0.0, 0.0])
    w0, x0, y0, z0 = x[3:7] / np.linalg.norm(x[3:7])
    w1, x1, y1, z1 = dq
    qn = np.array([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                   w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1])
    x[3:7] = qn / np.linalg.norm(qn)
    x[7:NQ] += rng.uniform(-0.05, 0.05, NQ - 7)
    if jnt_range is not None:
        lo, hi = jnt_range[:, 0], jnt_range[:, 1]
        m = 0.1 * (hi - lo)
        x[7:NQ] = np.clip(x[7:NQ], lo + m, hi - m)
    x[NQ:] += rng.uniform(-0.1, 0.1, NV)
    return x


def reference_set(tag, kinematics, com_velocity=None):
    d = np.load(os.path.join(ROOT, "data", "h1_refs.npz"))
    return ReferenceSet(d[f"{tag}_q"], d[f"{tag}_v"], d[f"{tag}_contact"], kinematics, com_velocity)


def walking_instances(ids, kinematics, N=N_HORIZON, refs=None):
    """Config 5 instances with GLOBAL ids `ids`: (window tuple of per-instance arrays, x0 [n][51], t0 [n])."""
    refs = refs if refs is not None else reference_set("walking", kinematics)
    ids = np.asarray(ids, dtype=np.int64)
    t0 = ids % (refs.T - (N + 1))
    lut = {int(t): refs.window(int(t), N) for t in np.unique(t0)}
    win = tuple(np.ascontiguousarray(np.stack([lut[int(t)][k] for t in t0])) for k in range(6))
    x0 = np.vstack([perturb(refs.x_ref_full[t], i) for t, i in zip(t0, ids)])
    return win, x0, t0


def standing_instances(ids, kinematics, N=N_HORIZON, jnt_range=None, refs=None):
    """Config 3 instances: one shared standing window, perturbed standing poses keyed by the global id."""
    refs = refs if refs is not None else reference_set("standing", kinematics)
    win = refs.window(0, N)
    x0 = np.vstack([perturb(standing_state(), i, jnt_range=jnt_range) for i in np.asarray(ids, dtype=np.int64)])
    return win, x0
