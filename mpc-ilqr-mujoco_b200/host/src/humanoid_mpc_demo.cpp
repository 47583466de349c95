// Closed-loop demo on the host API: the same sequence of calls the reference's main/humanoid_mpc.cpp makes
// (load config -> set up RobotUtils -> MPC -> loop getState / stepOnce / setControl / step), written for this
// repository. Usage: humanoid_mpc_demo [config.yaml] [sim_steps] [step_log.csv]
// With a third argument the wide step CSV (MPC::enableCSVLogging) is written there; q_optimal.csv / u_optimal.csv go to
// the configured results directory when save_trajectories is set. The last line is machine readable:
// BENCH_JSON {"step_ms": [...], "cost": [...]} — wall time of every MPC::stepOnce call (host buffers in and out).
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <vector>
#include "common/config.hpp"
#include "common/robot_utils.hpp"
#include "ilqr/mpc.hpp"

int main(int argc, char** argv) {
  Config config = loadConfigFromFile(argc > 1 ? argv[1] : "config.yaml");
  if (argc > 2) config.mpc.sim_steps = std::atoi(argv[2]);
  RobotUtils robot;
  if (!robot.loadModel(config.model_path)) return 1;
  robot.setContactImpratio(config.mpc.contact_impratio);
  robot.setTimeStep(config.mpc.physics_dt);
  robot.setGravity(config.mpc.gravity[0], config.mpc.gravity[1], config.mpc.gravity[2]);
  robot.initializeStandingPose();
  config.buildCostMatrices(robot.nx(), robot.nu(), robot.nq());
  robot.setCostWeights(config.Q, config.R, config.Qf);
  robot.setCoMWeight(config.mpc.costs.W_com); robot.setCoMVelWeight(config.mpc.costs.W_com_vel);
  robot.setEEPosWeight(config.mpc.costs.W_foot); robot.setEEVelWeight(config.mpc.costs.W_foot_vel);
  robot.setUprightWeight(config.mpc.costs.W_upright); robot.setBalanceWeight(config.mpc.costs.w_balance);
  robot.setConstraintWeights(config.mpc.joint_limit_weight, config.mpc.torque_limit_weight);
  if (!robot.loadReferences(config.q_ref_path, config.v_ref_path)) { std::cerr << "Failed to load reference trajectories." << std::endl; return 1; }
  if (!robot.loadContactSchedule(config.contact_schedule_path)) std::cerr << "Warning: no contact schedule" << std::endl;
  MPC mpc(robot, config.mpc.horizon, config.mpc.dt, config.urdf_path);
  if (argc > 3) mpc.enableCSVLogging(argv[3]);
  if (config.save_trajectories) mpc.enableOptimalTrajectoryLogging(config.results_path);
  std::vector<double> step_ms, step_cost;
  double total_ms = 0.0;
  for (int step = 0; step < config.mpc.sim_steps; ++step) {
    Eigen::VectorXd x(robot.nx()), u(robot.nu());
    robot.getState(x);
    if (!x.allFinite()) { std::cerr << "NaN detected in state at step " << step << ", breaking." << std::endl; break; }
    auto t0 = std::chrono::steady_clock::now();
    bool ok = mpc.stepOnce(x, u);
    step_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    total_ms += step_ms.back();
    step_cost.push_back(mpc.getLastSolveCost());
    if (!ok || !u.allFinite()) u.setZero();
    robot.setControl(u);
    robot.step();
    std::cout << "Step " << step << "/" << config.mpc.sim_steps << " | Cost: " << mpc.getLastSolveCost() << " | (X,Y,Z): (" << x(0)
              << "," << x(1) << "," << x(2) << ") m | Control range: [" << u.minCoeff() << ", " << u.maxCoeff() << "]" << std::endl;
  }
  std::cout << "Average MPC_stepOnce time: " << total_ms / std::max(1, config.mpc.sim_steps) << " ms" << std::endl;
  mpc.finalizeCSVLog();
  if (config.save_trajectories) mpc.finalizeOptimalTrajectoryLog();
  std::cout.precision(17);
  std::cout << "BENCH_JSON {\"step_ms\": [";
  for (size_t i = 0; i < step_ms.size(); ++i) std::cout << (i ? ", " : "") << step_ms[i];
  std::cout << "], \"cost\": [";
  for (size_t i = 0; i < step_cost.size(); ++i) std::cout << (i ? ", " : "") << step_cost[i];
  std::cout << "]}" << std::endl;
  return 0;
}
