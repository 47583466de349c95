// Run-time readers of the H1 model files into the plain-C H1Model description (include/h1_model.h):
//   load_mjcf_model  — the "dynamics" model from the MJCF that RobotUtils::loadModel opens
//                      (reference: src/common/robot_utils.cpp:19-55, mj_loadXML on robots/h1_description/mjcf/scene.xml,
//                      which includes h1.xml)
//   load_urdf_model  — the "cost" model from the URDF that symDerivatives builds with Pinocchio
//                      (reference: src/common/derivatives.cpp:26-39, pinocchio::urdf::buildModel with a free-flyer root)
// Both are small hand-written XML readers (the image has neither MuJoCo nor Pinocchio nor an XML library); they extract the
// tree topology, placements, inertial parameters, limits and actuator ranges — the same quantities tools/gen_h1_model.py
// writes into include/h1_model_data.h at build time — and reject anything outside the H1 model class (20 bodies, one
// axis-aligned hinge per body, DFS order). Contact-point and solver constants that are not part of the files (sole
// points, contact stiffness / damping, time step, gravity) are taken from `defaults`.
#pragma once
#include <string>
#include <vector>
#include "../../../../include/h1_model.h"

// Returns false and fills `error` when the file cannot be read or is not an H1-class model. `joint_names` receives the 19
// hinge names in dof order, `body_names` the 20 body names.
bool load_mjcf_model(const std::string& path, const H1Model& defaults, H1Model* out, std::vector<std::string>* joint_names,
                     std::vector<std::string>* body_names, std::string* error);
// The URDF joints are looked up by the MJCF's joint names so that both models agree on what qpos[7+i] means.
bool load_urdf_model(const std::string& path, const H1Model& defaults, const std::vector<std::string>& joint_names,
                     const std::vector<std::string>& body_names, H1Model* out, std::string* error);
