#!/usr/bin/env python3
"""Generate include/h1_model_data.h from the reference's H1 model files.

Reads (at generation time only, in the build container):
  /root/reference/robots/h1_description/mjcf/h1.xml   -> H1_DYNAMICS_MODEL  ("dynamics" model)
  /root/reference/robots/h1_description/urdf/h1.urdf  -> H1_COST_MODEL      ("cost" model)
and writes numeric tables only (tree topology, placements, inertial parameters, limits).

The joint order is keyed by joint NAME: the MJCF depth-first order defines dof order
 and the URDF joints are looked up by the same names, so the two
models are guaranteed to agree on what qpos[7+i] means.
"""
import sys
import xml.etree.ElementTree as ET
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = sys.argv[2] if len(sys.argv) > 2 else "include/h1_model_data.h"


def quat2mat(q):
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rpy2mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def vec(s, n=3):
    v = [float(t) for t in s.split()]
    assert len(v) == n, s
    return np.array(v)


def axis_index(a):
    a = np.asarray(a)
    idx = int(np.argmax(np.abs(a)))
    e = np.zeros(3)
    e[idx] = 1.0
    assert np.allclose(a, e), f"non axis-aligned joint axis {a}"
    return idx


def load_mjcf(path):
    root = ET.parse(path).getroot()
    dflt = root.find("default").find("default").find("joint")
    d_damp, d_arm = float(dflt.get("damping")), float(dflt.get("armature"))
    bodies = []

    def visit(elem, parent):
        idx = len(bodies)
        inert = elem.find("inertial")
        Ri = quat2mat(vec(inert.get("quat"), 4))
        I = Ri @ np.diag(vec(inert.get("diaginertia"))) @ Ri.T
        joint = elem.find("joint")
        b = dict(name=elem.get("name"), parent=parent,
                 pos=vec(elem.get("pos", "0 0 0")),
                 rfix=quat2mat(vec(elem.get("quat", "1 0 0 0"), 4)),
                 has_rfix=int(elem.get("quat") is not None),
                 mass=float(inert.get("mass")), ipos=vec(inert.get("pos")), I=I)
        if joint is not None:
            b.update(jname=joint.get("name"), axis=axis_index(vec(joint.get("axis"))),
                     range=vec(joint.get("range"), 2), damping=d_damp, armature=d_arm)
        else:
            assert elem.find("freejoint") is not None
            b.update(jname=None, axis=-1)
        bodies.append(b)
        for ch in elem.findall("body"):
            visit(ch, idx)

    visit(root.find("worldbody").find("body"), -1)
    bodies[0]["pos"] = np.zeros(3)  # the base placement is qpos[0:7]
    ctrl = {}
    for m in root.find("actuator").findall("motor"):
        ctrl[m.get("joint")] = vec(m.get("ctrlrange"), 2)
    # actuator order must equal hinge order (u[i] drives dof 6+i)
    order = [m.get("joint") for m in root.find("actuator").findall("motor")]
    assert order == [b["jname"] for b in bodies[1:]], "actuator order != joint order"
    for b in bodies[1:]:
        b["ctrl"] = ctrl[b["jname"]]
    return bodies


def load_urdf(path, joint_order, body_order):
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {j.get("name"): j for j in root.findall("joint")}
    # fixed-joint children must be massless (Pinocchio would otherwise merge their inertia)
    for j in joints.values():
        if j.get("type") == "fixed":
            assert links[j.find("child").get("link")].find("inertial") is None
    # URDF parser visits children sorted by joint name; verify that reproduces the MJCF order
    def sorted_dfs(link):
        out = []
        for jn in sorted(n for n, j in joints.items()
                         if j.find("parent").get("link") == link and j.get("type") != "fixed"):
            out.append(jn)
            out += sorted_dfs(joints[jn].find("child").get("link"))
        return out
    assert sorted_dfs("pelvis") == joint_order, "URDF sorted joint order != MJCF order"

    def inertial(link):
        it = links[link].find("inertial")
        o = it.find("origin")
        assert np.allclose(vec(o.get("rpy")), 0)
        i = it.find("inertia")
        g = lambda k: float(i.get(k))
        I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")],
                      [g("ixz"), g("iyz"), g("izz")]])
        return float(it.find("mass").get("value")), vec(o.get("xyz")), I

    m, c, I = inertial("pelvis")
    bodies = [dict(name="pelvis", parent=-1, pos=np.zeros(3), rfix=np.eye(3), has_rfix=0,
                   mass=m, ipos=c, I=I, axis=-1, jname=None)]
    for jn in joint_order:
        j = joints[jn]
        child = j.find("child").get("link")
        parent = j.find("parent").get("link")
        o = j.find("origin")
        rpy = vec(o.get("rpy"))
        m, c, I = inertial(child)
        lim = j.find("limit")
        bodies.append(dict(name=child, parent=body_order.index(parent), pos=vec(o.get("xyz")),
                           rfix=rpy2mat(rpy), has_rfix=int(np.any(rpy != 0)), mass=m, ipos=c, I=I,
                           axis=axis_index(vec(j.find("axis").get("xyz"))), jname=jn,
                           range=np.array([float(lim.get("lower")), float(lim.get("upper"))]),
                           ctrl=np.array([-float(lim.get("effort")), float(lim.get("effort"))]),
                           damping=0.0, armature=0.0))
    assert [b["name"] for b in bodies] == body_order
    return bodies


def fmt(x):
    return repr(float(x))


def arr(v):
    return "{" + ", ".join(fmt(x) for x in np.asarray(v).ravel()) + "}"


def emit(name, bodies, contact):
    nb = len(bodies)
    assert nb == 20
    L = [f"static const H1Model {name} = {{"]
    L.append("  /* parent */ {" + ", ".join(str(b["parent"]) for b in bodies) + "},")
    L.append("  /* axis */ {" + ", ".join(str(b["axis"]) for b in bodies) + "},")
    L.append("  /* has_rfix */ {" + ", ".join(str(b["has_rfix"]) for b in bodies) + "},")
    L.append("  /* pos */ {" + ",\n    ".join(arr(b["pos"]) for b in bodies) + "},")
    L.append("  /* rfix */ {" + ",\n    ".join(arr(b["rfix"]) for b in bodies) + "},")
    L.append("  /* mass */ " + arr([b["mass"] for b in bodies]) + ",")
    L.append("  /* ipos */ {" + ",\n    ".join(arr(b["ipos"]) for b in bodies) + "},")
    L.append("  /* inertia xx yy zz xy xz yz */ {" + ",\n    ".join(
        arr([b["I"][0, 0], b["I"][1, 1], b["I"][2, 2], b["I"][0, 1], b["I"][0, 2], b["I"][1, 2]])
        for b in bodies) + "},")
    L.append("  /* armature */ " + arr([0] * 6 + [b["armature"] for b in bodies[1:]]) + ",")
    L.append("  /* damping */ " + arr([0] * 6 + [b["damping"] for b in bodies[1:]]) + ",")
    L.append("  /* jnt_range */ {" + ", ".join(arr(b["range"]) for b in bodies[1:]) + "},")
    L.append("  /* ctrl_range */ {" + ", ".join(arr(b["ctrl"]) for b in bodies[1:]) + "},")
    names = [b["name"] for b in bodies]
    L.append(f"  /* foot_body */ {{{names.index('left_ankle_link')}, {names.index('right_ankle_link')}}},")
    pts = contact["pts"]
    L.append("  /* foot_pts */ {" + ", ".join("{" + ", ".join(arr(p) for p in pts) + "}" for _ in range(2)) + "},")
    L.append("  /* gravity */ " + arr(contact["gravity"]) + ",")
    L.append(f"  /* timestep */ {fmt(contact['timestep'])},")
    L.append(f"  /* contact kn bn bt eps */ {fmt(contact['kn'])}, {fmt(contact['bn'])}, {fmt(contact['bt'])}, {fmt(contact['eps'])},")
    L.append(f"  /* total_mass */ {fmt(sum(b['mass'] for b in bodies))}")
    L.append("};")
    return "\n".join(L)


def main():
    mj = load_mjcf(f"{REF}/robots/h1_description/mjcf/h1.xml")
    order = [b["jname"] for b in mj[1:]]
    names = [b["name"] for b in mj]
    ur = load_urdf(f"{REF}/robots/h1_description/urdf/h1.urdf", order, names)
    # Sole contact points: corners of the convex hull of {left,right}_ankle_link.STL measured in the
    # ankle frame (x in [-0.1, 0.2], y in [-0.015, 0.015], sole plane z = -0.07).
    contact = dict(pts=[[-0.1, -0.015, -0.07], [-0.1, 0.015, -0.07], [0.2, -0.015, -0.07], [0.2, 0.015, -0.07]],
                   gravity=[0.0, 0.0, -1.0],  # config.yaml:20 as shipped
                   timestep=0.02, kn=3.0e4, bn=1.0e3, bt=1.0e3, eps=2.0e-3)
    with open(OUT, "w") as f:
        f.write("/* GENERATED by tools/gen_h1_model.py from the reference's h1.xml / h1.urdf — do not edit. */\n")
        f.write("#ifndef H1_MODEL_DATA_H\n#define H1_MODEL_DATA_H\n#include \"h1_model.h\"\n\n")
        f.write("/* body order: " + " ".join(names) + " */\n")
        f.write(emit("H1_DYNAMICS_MODEL", mj, contact) + "\n\n")
        f.write(emit("H1_COST_MODEL", ur, contact) + "\n\n#endif\n")
    print("wrote", OUT, "mass mjcf", sum(b["mass"] for b in mj), "urdf", sum(b["mass"] for b in ur))


if __name__ == "__main__":
    main()
