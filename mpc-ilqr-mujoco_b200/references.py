"""Reference-trajectory inputs of the hot path (host side).

Mirrors RobotUtils::loadReferences / loadContactSchedule / getReferenceWindow / isStance / getEEReference
(/root/reference/src/common/robot_utils.cpp:281-443, 445-504, 525-532):
  - q / v CSVs are read in lockstep, rows whose column counts are not (26, 25) are skipped;
  - the contact CSV has one header line, then `int,int` rows (1 = stance);
  - per reference row the CoM and the two ankle-body positions are computed on the DYNAMICS (MJCF) model;
  - windows clamp at the last row; the contact flags / foot targets / CoM-velocity targets are looked up with
    the HORIZON-LOCAL knot index (reference quirk Q6), unless `schedule_offset=True` is requested.
"""
import numpy as np

from .ctypes_defs import NQ, NV, NX, NU


def load_qv_csv(q_path, v_path):
    qs, vs = [], []
    with open(q_path) as fq, open(v_path) as fv:
        for ql, vl in zip(fq, fv):
            def parse(line):
                vals = []
                for tok in line.strip().split(","):
                    try:
                        vals.append(float(tok))
                    except ValueError:
                        continue
                return vals
            q, v = parse(ql), parse(vl)
            if len(q) != NQ or len(v) != NV:
                continue
            qs.append(q)
            vs.append(v)
    if not qs:
        raise RuntimeError("No valid reference states loaded")
    return np.asarray(qs, dtype=np.float64), np.asarray(vs, dtype=np.float64)


def pinocchio_to_mujoco_q(q_pin):
    """Configuration rows in Pinocchio order [x y z | qx qy qz qw | joints] -> MuJoCo order [x y z | qw qx qy qz | joints]
    (the conversion of the reference's tooling, get_contacts.py:18-41; data/h1_walking_pin.csv, the file BASELINE config 2
    names, is stored in Pinocchio order). Works on one row or a [T][26] array."""
    q = np.array(q_pin, dtype=np.float64, copy=True)
    src = np.asarray(q_pin, dtype=np.float64)
    q[..., 3] = src[..., 6]
    q[..., 4] = src[..., 3]
    q[..., 5] = src[..., 4]
    q[..., 6] = src[..., 5]
    return q


def load_q_pin_csv(q_pin_path):
    """A Pinocchio-ordered configuration CSV (h1_walking_pin.csv) as MuJoCo-ordered rows [T][26]; rows whose column
    count is not 26 are skipped like loadReferences does. The file carries no velocities: pair it with a v CSV through
    load_qv_csv semantics, or use `finite_difference_velocities` for rows beyond the shipped v_ref2.csv (400 rows)."""
    rows = []
    with open(q_pin_path) as f:
        for line in f:
            vals = []
            for tok in line.strip().split(","):
                try:
                    vals.append(float(tok))
                except ValueError:
                    continue
            if len(vals) == NQ:
                rows.append(vals)
    if not rows:
        raise RuntimeError("No valid reference states loaded")
    return pinocchio_to_mujoco_q(np.asarray(rows, dtype=np.float64))


def load_contact_csv(path):
    rows = []
    with open(path) as f:
        f.readline()  # header
        for line in f:
            vals = []
            for tok in line.strip().split(","):
                try:
                    vals.append(int(tok))
                except ValueError:
                    continue
            if vals:
                rows.append(vals)
    return np.asarray(rows, dtype=np.int32)


def contact_schedule_from_states(q, sole_points, threshold=1e-3):
    """Contact schedule [T][2] (left, right; 1 = stance) of a reference motion q [T][26]: the counterpart of the
    reference's offline tool (/root/reference/get_contacts.py:96-157), which runs MuJoCo's collision detection on
    the foot geoms and marks a foot in contact when a contact has dist < 1e-3. Here a foot is in contact when
    one of its four sole points (the contact points of f_D) is lower than `threshold` above the ground plane.
    `sole_points(x[n,51]) -> [n][8][3]` is H1IlqrBatch.sole_points (GPU)."""
    q = np.asarray(q, dtype=np.float64)
    x = np.hstack([q, np.zeros((q.shape[0], NV))])
    z = np.asarray(sole_points(x))[:, :, 2].reshape(q.shape[0], 2, 4)
    return (z.min(axis=2) < threshold).astype(np.int32)


def write_contact_csv(path, contact):
    """Same file format as get_contacts.py writes (pandas to_csv, index=False): header + `int,int` rows."""
    with open(path, "w") as f:
        f.write("left_foot,right_foot\n")
        for l, r in np.asarray(contact, dtype=int):
            f.write(f"{l},{r}\n")


class ReferenceSet:
    """x_ref_full / com_ref_full / ee_pos_ref_full / contact schedule of one reference motion."""

    def __init__(self, q, v, contact, kinematics, com_velocity=None):
        """`kinematics(x[n,51]) -> (com[n,3], ee[n,2,3])` on the dynamics model; the product passes
        H1IlqrBatch.reference_kinematics (GPU), the tests may pass the oracle's. `com_velocity(x[n,51]) -> [n,3]`
        is the per-row CoM-velocity target J_subtreeCom * qvel of loadReferences (robot_utils.cpp:388-397;
        H1IlqrBatch.reference_com_velocity on the GPU). It is only tracked when W_com_vel > 0 (0 as shipped); without
        the callable the targets are zero and `require_com_velocity` refuses a positive weight."""
        self.x_ref_full = np.ascontiguousarray(np.hstack([q, v]))
        self.T = self.x_ref_full.shape[0]
        self.com_ref_full, self.ee_pos_ref_full = kinematics(self.x_ref_full)
        self.contact = np.asarray(contact, dtype=np.int32)
        self.has_com_velocity = com_velocity is not None
        self.com_vel_ref_full = (np.ascontiguousarray(com_velocity(self.x_ref_full)) if com_velocity is not None
                                 else np.zeros((self.T, 3)))

    def require_com_velocity(self, weights):
        """Raise instead of silently tracking a zero CoM velocity when the weight is positive."""
        if weights.w_com_vel > 0.0 and not self.has_com_velocity:
            raise RuntimeError("W_com_vel > 0 needs the CoM-velocity targets: build the ReferenceSet with "
                               "com_velocity=H1IlqrBatch.reference_com_velocity")

    def is_stance(self, ee, t):
        if t < 0 or t >= self.contact.shape[0] or ee < 0 or ee >= self.contact.shape[1]:
            return 1
        return int(self.contact[t, ee] == 1)

    def window(self, t0, N, schedule_offset=False):
        """(x_ref[N+1,51], u_ref[N,19], com_ref[N+1,3], ee_ref[N+1,2,3], stance[N+1,2], com_vel_ref[N+1,3])"""
        idx = np.minimum(t0 + np.arange(N + 1), self.T - 1)
        x_ref = self.x_ref_full[idx]
        u_ref = np.zeros((N, NU))
        com_ref = self.com_ref_full[idx]
        loc = np.arange(N + 1) + (t0 if schedule_offset else 0)
        if loc.max() >= self.T:
            if not schedule_offset:
                raise RuntimeError("Invalid reference index")  # getEEReference throws (robot_utils.cpp:526-529)
            loc = np.minimum(loc, self.T - 1)
        ee_ref = self.ee_pos_ref_full[loc]
        stance = np.array([[self.is_stance(e, int(t)) for e in range(2)] for t in loc], dtype=np.int32)
        return (np.ascontiguousarray(x_ref), u_ref, np.ascontiguousarray(com_ref), np.ascontiguousarray(ee_ref),
                stance, np.ascontiguousarray(self.com_vel_ref_full[loc]))


def standing_state():
    """RobotUtils::initializeStandingPose (robot_utils.cpp:557-596)."""
    x = np.zeros(NX)
    x[2] = 1.0432
    x[3] = 1.0
    return x


def perturbed_states(x_nominal, batch, seed=0, jnt_range=None):
    """SURVEY.md §8(d) config 3/5 perturbation: base xyz U(+-0.02), orientation exp(U(+-0.05)^3), joints
    U(+-0.05) clipped to the inner 80% of range, velocities U(+-0.1). Counter-based (Philox) per instance."""
    x_nominal = np.atleast_2d(np.asarray(x_nominal, dtype=np.float64))
    out = np.empty((batch, NX))
    for i in range(batch):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[i, 0, 0, 0]))
        x = x_nominal[i % x_nominal.shape[0]].copy()
        x[0:3] += rng.uniform(-0.02, 0.02, 3)
        rv = rng.uniform(-0.05, 0.05, 3)
        ang = np.linalg.norm(rv)
        dq = np.array([np.cos(ang / 2), *(np.sin(ang / 2) / ang * rv)]) if ang > 0 else np.array([1.0, 0, 0, 0])
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
        out[i] = x
    return out
