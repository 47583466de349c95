// Exercises the RobotUtils entry points that main/humanoid_mpc.cpp does not call (reference: include/common/
// robot_utils.hpp:51-57, 74-81, 95-107) and prints their results as one JSON object; tests/test_host_cpp_gpu.py compares
// them with the CPU oracle. Usage: host_api_check config.yaml
#include <iostream>
#include <vector>
#include "common/config.hpp"
#include "common/robot_utils.hpp"

static void arr(const char* name, const double* v, long n, bool last = false) {
  std::cout << "\"" << name << "\": [";
  for (long i = 0; i < n; ++i) std::cout << (i ? ", " : "") << v[i];
  std::cout << "]" << (last ? "" : ", ");
}

int main(int argc, char** argv) {
  Config config = loadConfigFromFile(argc > 1 ? argv[1] : "config.yaml");
  RobotUtils robot;
  if (!robot.loadModel(config.model_path)) return 1;
  robot.setTimeStep(config.mpc.physics_dt);
  robot.setGravity(config.mpc.gravity[0], config.mpc.gravity[1], config.mpc.gravity[2]);
  robot.initializeStandingPose();
  config.buildCostMatrices(robot.nx(), robot.nu(), robot.nq());
  robot.setCostWeights(config.Q, config.R, config.Qf);
  robot.setCoMWeight(config.mpc.costs.W_com);
  robot.setConstraintWeights(config.mpc.joint_limit_weight, config.mpc.torque_limit_weight);
  if (!robot.loadReferences(config.q_ref_path, config.v_ref_path)) return 1;
  robot.loadContactSchedule(config.contact_schedule_path);
  const int nx = robot.nx(), nu = robot.nu(), T = 37;
  robot.resetToReference(T);
  Eigen::VectorXd x(nx), u(nu);
  robot.getState(x);
  for (int i = 0; i < nu; ++i) { x(7 + i) += 0.35 * ((i % 3) - 1) + 0.02 * i; u(i) = 40.0 * ((i % 5) - 2) + 3.0 * i; }
  std::cout.precision(17);
  std::cerr.precision(17);
  std::streambuf* keep = std::cout.rdbuf(std::cerr.rdbuf());   // the library's progress messages go to stderr from here on
  Eigen::MatrixXd A, B, Hxx, Huu;
  robot.linearizeDynamicsFD(x, u, A, B);
  Eigen::VectorXd gx, gu;
  robot.constraintGradients(x, u, gx, gu);
  robot.constraintHessians(x, u, Hxx, Huu);
  std::vector<double> hx(nx), hu(nu);
  for (int i = 0; i < nx; ++i) hx[i] = Hxx(i, i);
  for (int i = 0; i < nu; ++i) hu[i] = Huu(i, i);
  const double cc = robot.constraintCost(x, u), sc = robot.stageCost(T, x, u), sc_far = robot.stageCost(100000, x, u), tc = robot.terminalCost(x);
  Eigen::Vector3d ev0 = robot.getEEVelReference(T, 0), ev1 = robot.getEEVelReference(T, 1), cv = robot.getCoMVelReference(T);
  // whole symmetric Q / R / Qf (off-diagonal entries) through setCostWeights
  Eigen::MatrixXd Qs = config.Q, Rs = config.R, Qfs = config.Qf;
  Qs(0, 1) = Qs(1, 0) = 7.0; Qs(30, 8) = Qs(8, 30) = -3.0; Rs(2, 5) = Rs(5, 2) = 0.0004; Qfs(10, 40) = Qfs(40, 10) = 11.0;
  robot.setCostWeights(Qs, Rs, Qfs);
  const double sc_full = robot.stageCost(T, x, u), tc_full = robot.terminalCost(x);
  robot.setCostWeights(config.Q, config.R, config.Qf);
  const double sc_back = robot.stageCost(T, x, u);
  Eigen::VectorXd xn(nx), xn_heavy(nx);
  robot.rolloutOneStep(x, u, xn);
  robot.scaleRobotMass(1.5);
  robot.rolloutOneStep(x, u, xn_heavy);
  std::cout.rdbuf(keep);
  std::cout << "{";
  arr("x", x.data(), nx); arr("u", u.data(), nu); arr("A", A.data(), (long)nx * nx); arr("B", B.data(), (long)nx * nu);
  arr("grad_x", gx.data(), nx); arr("grad_u", gu.data(), nu); arr("hess_xx_diag", hx.data(), nx); arr("hess_uu_diag", hu.data(), nu);
  arr("ee_vel_0", ev0.data(), 3); arr("ee_vel_1", ev1.data(), 3); arr("com_vel", cv.data(), 3);
  arr("x_next", xn.data(), nx); arr("x_next_heavy", xn_heavy.data(), nx);
  std::cout << "\"constraint_cost\": " << cc << ", \"stage_cost\": " << sc << ", \"stage_cost_far\": " << sc_far << ", \"terminal_cost\": " << tc << ", \"stage_cost_fullq\": " << sc_full << ", \"terminal_cost_fullq\": " << tc_full
            << ", \"stage_cost_back\": " << sc_back
            << ", \"joint_id_torso\": " << robot.jointId("torso_joint") << ", \"joint_id_left_knee\": " << robot.jointId("left_knee_joint")
            << ", \"joint_id_unknown\": " << robot.jointId("nope") << ", \"t\": " << T << "}" << std::endl;
  return 0;
}
