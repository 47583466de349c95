// iLQR with the reference's public API (reference: include/ilqr/ilqr.hpp:17-45). Every stage runs on the GPU
// through the C ABI; this class only marshals Eigen-style containers and mirrors the accessors.
#pragma once
#include <vector>
#include "common/robot_utils.hpp"

class iLQR {
 public:
  iLQR(RobotUtils& robot, int N, double dt, const std::string& urdf_path);
  ~iLQR();
  iLQR(const iLQR&) = delete;
  iLQR& operator=(const iLQR&) = delete;

  void setRegularization(double lambda);
  void setMaxIterations(int max_iter);
  void setTolerance(double tol);

  bool solve(const Eigen::VectorXd& x0, const std::vector<Eigen::VectorXd>& x_ref,
             const std::vector<Eigen::VectorXd>& u_ref, const std::vector<Eigen::Vector3d>& com_ref, double& cost_out);

  const std::vector<Eigen::VectorXd>& xbar() const { return xbar_; }
  const std::vector<Eigen::VectorXd>& ubar() const { return ubar_; }
  const std::vector<Eigen::MatrixXd>& gainsK() const { return K_; }
  const std::vector<Eigen::VectorXd>& gainsKff() const { return kff_; }

  void initializeWithReference(const Eigen::VectorXd& x0, const std::vector<Eigen::VectorXd>& x_ref,
                               const std::vector<Eigen::VectorXd>& u_ref, const std::vector<Eigen::Vector3d>& com_ref,
                               const std::vector<Eigen::VectorXd>* prev_xbar = nullptr,
                               const std::vector<Eigen::VectorXd>* prev_ubar = nullptr);

  H1StageTimes lastStageTimes() const;   // CUDA-event timings under the reference's profiling labels

 private:
  bool recreate();
  void sync_model();
  bool upload_window(const std::vector<Eigen::VectorXd>& x_ref, const std::vector<Eigen::VectorXd>& u_ref,
                     const std::vector<Eigen::Vector3d>& com_ref);
  void download_solution();
  RobotUtils& robot_;
  int N_;
  double dt_;
  H1SolverOptions opt_;
  H1Ilqr* h_;
  int model_version_ = -1;
  H1Model cost_model_;
  bool cost_from_file_ = false;
  std::vector<Eigen::VectorXd> xbar_, ubar_, kff_;
  std::vector<Eigen::MatrixXd> K_;
};
