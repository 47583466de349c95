#!/usr/bin/env python3
"""Pack the reference's CSV inputs that the hot path consumes into data/h1_refs.npz (inputs only).

Source files (read at generation time, in the build container only):
  /root/reference/data/q_ref2_mj.csv, v_ref2.csv, contact_walking.csv      (shipped default, config.yaml:12-14)
  /root/reference/data/q_standing.csv, v_standing.csv, contact_standing.csv (BASELINE config 1)
The loaders in mpc-ilqr-mujoco_b200/references.py apply the reference's parsing rules
(src/common/robot_utils.cpp:281-347, 445-492) to the CSV text; this script just runs them.
"""
import sys
import numpy as np
sys.path.insert(0, ".")
from mpc_ilqr_mujoco_b200.references import load_contact_csv, load_qv_csv

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {}
for tag, q, v, c in (("walking", "q_ref2_mj.csv", "v_ref2.csv", "contact_walking.csv"),
                     ("standing", "q_standing.csv", "v_standing.csv", "contact_standing.csv")):
    Q, V = load_qv_csv(f"{REF}/data/{q}", f"{REF}/data/{v}")
    out[f"{tag}_q"], out[f"{tag}_v"] = Q, V
    out[f"{tag}_contact"] = load_contact_csv(f"{REF}/data/{c}")
    print(tag, Q.shape, V.shape, out[f"{tag}_contact"].shape)
# first rows of the Pinocchio-ordered walking file (BASELINE config 2 names it): fixture of the order conversion test
import csv
with open(f"{REF}/data/h1_walking_pin.csv") as f:
    out["walking_pin_q_head"] = np.array([[float(t) for t in row] for _, row in zip(range(64), csv.reader(f))])
np.savez_compressed("data/h1_refs.npz", **out)
# the reference's model files (robots/h1_description/mjcf/{scene,h1}.xml, urdf/h1.urdf) as byte arrays: inputs of the run-time
# model loader test (host/src/model_loader.cpp) and of bench / demo runs that want RobotUtils::loadModel to parse real files
models = {}
for key, rel in (("mjcf_scene_xml", "robots/h1_description/mjcf/scene.xml"), ("mjcf_h1_xml", "robots/h1_description/mjcf/h1.xml"),
                 ("urdf_h1_urdf", "robots/h1_description/urdf/h1.urdf")):
    models[key] = np.frombuffer(open(f"{REF}/{rel}", "rb").read(), dtype=np.uint8)
np.savez_compressed("data/h1_models.npz", **models)
