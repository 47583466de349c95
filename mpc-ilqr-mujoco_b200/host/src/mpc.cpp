// MPC orchestrator (reference: src/ilqr/mpc.cpp:16-179, 270-355): reference window -> warm start -> solve ->
// TV-LQR first control -> keep the solution for the next warm start; q_optimal.csv / u_optimal.csv logging in
// the reference's format.
#include "ilqr/mpc.hpp"
#include <chrono>
#include <iostream>

MPC::MPC(RobotUtils& robot, int N, double dt, const std::string& urdf_path)
    : robot_(robot), ilqr_(robot, N, dt, urdf_path), N_(N), dt_(dt), t_idx_(0), has_prev_solution_(false),
      last_solve_cost_(0.0), last_solve_time_ms_(0.0) {
  const int nx = robot_.nx(), nu = robot_.nu();
  x_ref_window_.assign(N_ + 1, Eigen::VectorXd::Zero(nx));
  u_ref_window_.assign(N_, Eigen::VectorXd::Zero(nu));
  prev_xbar_.assign(N_ + 1, Eigen::VectorXd::Zero(nx));
  prev_ubar_.assign(N_, Eigen::VectorXd::Zero(nu));
  prev_K_.assign(N_, Eigen::MatrixXd::Zero(nu, nx));
  std::cout << "MPC initialized with N=" << N_ << ", dt=" << dt_ << std::endl;
}

void MPC::extractReferenceWindow() { robot_.getReferenceWindow(t_idx_, N_, x_ref_window_, u_ref_window_, com_ref_window_); }

bool MPC::stepOnce(const Eigen::VectorXd& x_measured, Eigen::VectorXd& u_apply) {
  auto start = std::chrono::steady_clock::now();
  try {
    extractReferenceWindow();
    if (has_prev_solution_) ilqr_.initializeWithReference(x_measured, x_ref_window_, u_ref_window_, com_ref_window_, &prev_xbar_, &prev_ubar_);
    else ilqr_.initializeWithReference(x_measured, x_ref_window_, u_ref_window_, com_ref_window_);
    double solve_cost = 0.0;
    if (!ilqr_.solve(x_measured, x_ref_window_, u_ref_window_, com_ref_window_, solve_cost)) {
      std::cerr << "iLQR solve failed at time index " << t_idx_ << std::endl;
      u_apply = has_prev_solution_ ? prev_ubar_[0] : Eigen::VectorXd::Zero(robot_.nu());
      return false;
    }
    const auto& xbar = ilqr_.xbar(); const auto& ubar = ilqr_.ubar(); const auto& K = ilqr_.gainsK();
    u_apply = ubar[0] + K[0] * (x_measured - xbar[0]);
    prev_xbar_ = xbar; prev_ubar_ = ubar; prev_K_ = K;
    has_prev_solution_ = true;
    last_solve_cost_ = solve_cost;
    last_solve_time_ms_ = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - start).count() / 1000.0;
    t_idx_++;
    logCurrentStep(x_measured, u_apply);
    logAppliedOptimal(x_measured, u_apply);
    return true;
  } catch (const std::exception& e) {
    std::cerr << "Exception in MPC step: " << e.what() << std::endl;
    u_apply = Eigen::VectorXd::Zero(robot_.nu());
    return false;
  }
}

void MPC::reset() {
  t_idx_ = 0; has_prev_solution_ = false; last_solve_cost_ = 0.0; last_solve_time_ms_ = 0.0;
  for (auto& v : prev_xbar_) v.setZero();
  for (auto& v : prev_ubar_) v.setZero();
  for (auto& m : prev_K_) m.setZero();
  std::cout << "MPC reset" << std::endl;
}

void MPC::getNominalTrajectory(std::vector<Eigen::VectorXd>& x_traj, std::vector<Eigen::VectorXd>& u_traj) const {
  if (has_prev_solution_) { x_traj = prev_xbar_; u_traj = prev_ubar_; } else { x_traj.clear(); u_traj.clear(); }
}

// ---- wide step CSV (mpc.cpp:181-268): one row per MPC step, default ostream number formatting, std::endl ----
void MPC::enableCSVLogging(const std::string& filename) {
  csv_filename_ = filename;
  csv_file_.open(csv_filename_, std::ios::out | std::ios::trunc);
  if (!csv_file_.is_open()) { std::cerr << "Failed to open CSV file: " << csv_filename_ << std::endl; return; }
  csv_file_ << "time_index,time_sec,solve_cost,solve_time_ms";
  for (int i = 0; i < robot_.nx(); ++i) csv_file_ << ",x_" << i;
  for (int i = 0; i < robot_.nu(); ++i) csv_file_ << ",u_" << i;
  for (int i = 0; i < robot_.nx(); ++i) csv_file_ << ",x_ref_" << i;
  for (int i = 0; i < robot_.nu(); ++i) csv_file_ << ",u_ref_" << i;
  csv_file_ << std::endl;
  std::cout << "CSV logging started: " << csv_filename_ << std::endl;
}
void MPC::logCurrentStep(const Eigen::VectorXd& x_measured, const Eigen::VectorXd& u_applied) {
  if (!csv_file_.is_open()) return;
  csv_file_ << t_idx_ << "," << (t_idx_ * dt_) << "," << last_solve_cost_ << "," << last_solve_time_ms_;
  for (long i = 0; i < x_measured.size(); ++i) csv_file_ << "," << x_measured(i);
  for (long i = 0; i < u_applied.size(); ++i) csv_file_ << "," << u_applied(i);
  if (!x_ref_window_.empty()) { for (long i = 0; i < x_ref_window_[0].size(); ++i) csv_file_ << "," << x_ref_window_[0](i); }
  else { for (int i = 0; i < robot_.nx(); ++i) csv_file_ << ",0.0"; }
  if (!u_ref_window_.empty()) { for (long i = 0; i < u_ref_window_[0].size(); ++i) csv_file_ << "," << u_ref_window_[0](i); }
  else { for (int i = 0; i < robot_.nu(); ++i) csv_file_ << ",0.0"; }
  csv_file_ << std::endl;
}
void MPC::finalizeCSVLog() {
  if (csv_file_.is_open()) {
    csv_file_.flush();
    csv_file_.close();
    std::cout << "CSV log finalized: " << csv_filename_ << std::endl;
  }
}

// ---- q_optimal.csv / u_optimal.csv (mpc.cpp:270-355): first knot of the optimised trajectory per step, the format
//      simulate.py / plotter.py read; default ostream number formatting like the reference ----
void MPC::enableOptimalTrajectoryLogging(const std::string& base_path) {
  trajectory_base_path_ = base_path;
  q_optimal_file_.open(base_path + "/q_optimal.csv", std::ios::out | std::ios::trunc);
  u_optimal_file_.open(base_path + "/u_optimal.csv", std::ios::out | std::ios::trunc);
  if (!q_optimal_file_.is_open() || !u_optimal_file_.is_open()) {
    std::cerr << "Failed to open optimal trajectory files in: " << base_path << std::endl;
    return;
  }
  q_optimal_file_ << "step,time_sec";
  for (int i = 0; i < robot_.nq(); ++i) q_optimal_file_ << ",q_" << i;
  q_optimal_file_ << std::endl;
  u_optimal_file_ << "step,time_sec";
  for (int i = 0; i < robot_.nu(); ++i) u_optimal_file_ << ",u_" << i;
  u_optimal_file_ << std::endl;
}
void MPC::logAppliedOptimal(const Eigen::VectorXd& x_applied, const Eigen::VectorXd& u_applied) {
  if (!q_optimal_file_.is_open() || !u_optimal_file_.is_open()) return;
  const auto& x_optimal = ilqr_.xbar(); const auto& u_optimal = ilqr_.ubar();
  q_optimal_file_ << t_idx_ << "," << (t_idx_ * dt_);
  for (int i = 0; i < robot_.nq(); ++i) q_optimal_file_ << "," << (!x_optimal.empty() ? x_optimal[0](i) : x_applied(i));
  q_optimal_file_ << std::endl;
  u_optimal_file_ << t_idx_ << "," << (t_idx_ * dt_);
  if (!u_optimal.empty()) { for (long i = 0; i < u_optimal[0].size(); ++i) u_optimal_file_ << "," << u_optimal[0](i); }
  else { for (long i = 0; i < u_applied.size(); ++i) u_optimal_file_ << "," << u_applied(i); }
  u_optimal_file_ << std::endl;
}
void MPC::finalizeOptimalTrajectoryLog() {
  if (q_optimal_file_.is_open()) { q_optimal_file_.flush(); q_optimal_file_.close(); }
  if (u_optimal_file_.is_open()) { u_optimal_file_.flush(); u_optimal_file_.close(); }
  std::cout << "Optimal trajectory logs finalized: " << trajectory_base_path_ << "/q_optimal.csv and u_optimal.csv" << std::endl;
}
