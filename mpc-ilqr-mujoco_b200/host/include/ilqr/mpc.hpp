// MPC orchestrator with the reference's public API (reference: include/ilqr/mpc.hpp:18-81, src/ilqr/mpc.cpp).
#pragma once
#include <fstream>
#include <string>
#include <vector>
#include "ilqr/ilqr.hpp"

class MPC {
 public:
  MPC(RobotUtils& robot, int N, double dt, const std::string& urdf_path);
  bool stepOnce(const Eigen::VectorXd& x_measured, Eigen::VectorXd& u_apply);
  void reset();
  void setTimeIndex(int t_idx) { t_idx_ = t_idx; }
  int getTimeIndex() const { return t_idx_; }
  void enableCSVLogging(const std::string& filename);
  void logCurrentStep(const Eigen::VectorXd& x_measured, const Eigen::VectorXd& u_applied);
  void finalizeCSVLog();
  void enableOptimalTrajectoryLogging(const std::string& base_path);
  void logAppliedOptimal(const Eigen::VectorXd& x_applied, const Eigen::VectorXd& u_applied);
  void finalizeOptimalTrajectoryLog();
  const iLQR& solver() const { return ilqr_; }
  const std::vector<Eigen::MatrixXd>& gainsK() const { return prev_K_; }
  double getLastSolveCost() const { return last_solve_cost_; }
  double getLastSolveTimeMs() const { return last_solve_time_ms_; }
  void getNominalTrajectory(std::vector<Eigen::VectorXd>& x_traj, std::vector<Eigen::VectorXd>& u_traj) const;

 private:
  void extractReferenceWindow();
  RobotUtils& robot_;
  iLQR ilqr_;
  int N_;
  double dt_;
  int t_idx_;
  std::vector<Eigen::VectorXd> x_ref_window_, u_ref_window_;
  std::vector<Eigen::Vector3d> com_ref_window_;
  bool has_prev_solution_;
  std::vector<Eigen::VectorXd> prev_xbar_, prev_ubar_;
  std::vector<Eigen::MatrixXd> prev_K_;
  double last_solve_cost_, last_solve_time_ms_;
  std::string trajectory_base_path_;
  std::ofstream q_optimal_file_, u_optimal_file_;
  std::ofstream csv_file_;
  std::string csv_filename_;
};
