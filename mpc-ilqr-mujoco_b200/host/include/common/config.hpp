// Config / loadConfigFromFile / Config::buildCostMatrices with the reference's names and semantics
// (reference: include/common/config.hpp:9-54, src/common/config.cpp:4-122). Host-only.
#pragma once
#include <Eigen/Dense>
#include <string>
#include <vector>

struct CostWeights {
  double Q_position_x, Q_position_y, Q_position_z, Q_quat_w;
  std::vector<double> Q_quat_xyz;
  double Q_joint_pos, Q_vel_x, Q_vel_y, Q_vel_z, Q_ang_vel, Q_joint_vel;
  double R_control;
  double Qf_multiplier, Qf_position_x, Qf_position_y, Qf_position_z, Qf_vel_z;
  double W_com, W_com_vel, W_foot, W_foot_vel;
  double W_upright;
  double w_balance;
};

struct MpcParams {
  int horizon;
  double dt, physics_dt;
  std::vector<double> gravity;
  int sim_steps;
  double contact_impratio;
  CostWeights costs;
  double joint_limit_weight;
  double torque_limit_weight;
};

struct Config {
  std::string model_path, urdf_path, q_ref_path, v_ref_path, contact_schedule_path, results_path;
  bool verbose;
  bool save_trajectories;
  MpcParams mpc;
  Eigen::MatrixXd Q, R, Qf;
  void buildCostMatrices(int nx, int nu, int nq);
};

Config loadConfigFromFile(const std::string& filepath);
