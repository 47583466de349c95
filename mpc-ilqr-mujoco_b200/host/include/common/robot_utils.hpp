// RobotUtils with the reference's public API (reference: include/common/robot_utils.hpp:18-159), backed by the
// B200 solver core through the C ABI (include/h1ilqr.h) instead of MuJoCo. It owns the plant state (qpos/qvel/
// ctrl), the cost weights, the reference trajectory tables and the contact schedule; dynamics, bias forces and
// the per-row reference FK run on the GPU.
#pragma once
#include <mujoco/mujoco.h>
#include <Eigen/Dense>
#include <string>
#include <vector>
#include "../../../../include/h1ilqr.h"

class RobotUtils {
 public:
  RobotUtils();
  ~RobotUtils();
  RobotUtils(const RobotUtils&) = delete;
  RobotUtils& operator=(const RobotUtils&) = delete;

  bool loadModel(const std::string& xml_path);
  void setContactImpratio(double impratio);
  void setTimeStep(double dt);
  void setGravity(double gx = 0.0, double gy = 0.0, double gz = 0.0);

  int nx() const { return nx_; }
  int nu() const { return nu_; }
  int nq() const { return loaded_ ? H1_NQ : 0; }
  int nv() const { return loaded_ ? H1_NV : 0; }
  double dt() const { return dt_; }

  const Eigen::MatrixXd& Q() const { return Q_; }
  const Eigen::MatrixXd& R() const { return R_; }
  const Eigen::MatrixXd& Qf() const { return Qf_; }

  void setState(const Eigen::VectorXd& x);
  void getState(Eigen::VectorXd& x) const;
  void setControl(const Eigen::VectorXd& u);
  void step();
  void rolloutOneStep(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::VectorXd& x_next);
  void linearizeDynamicsFD(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::MatrixXd& A, Eigen::MatrixXd& B,
                           double eps = 1e-5);

  double stageCost(int t, const Eigen::VectorXd& x, const Eigen::VectorXd& u) const;
  double terminalCost(const Eigen::VectorXd& x) const;
  double constraintCost(const Eigen::VectorXd& x, const Eigen::VectorXd& u) const;
  void constraintGradients(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::VectorXd& grad_x,
                           Eigen::VectorXd& grad_u) const;
  void constraintHessians(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::MatrixXd& hess_xx,
                          Eigen::MatrixXd& hess_uu) const;

  void setCostWeights(const Eigen::MatrixXd& Q, const Eigen::MatrixXd& R, const Eigen::MatrixXd& Qf);
  void setCoMWeight(double w) { w_com_ = w; }
  double getCoMWeight() const { return w_com_; }
  void setCoMVelWeight(double w) { w_com_vel_ = w; }
  double getCoMVelWeight() const { return w_com_vel_; }
  void setEEPosWeight(double w) { w_ee_pos_ = w; }
  double getEEPosWeight() const { return w_ee_pos_; }
  void setEEVelWeight(double w) { w_ee_vel_ = w; }
  double getEEVelWeight() const { return w_ee_vel_; }
  void setUprightWeight(double w) { w_upright_ = w; }
  double getUprightWeight() const { return w_upright_; }
  void setBalanceWeight(double w) { w_balance_ = w; }
  double getBalanceWeight() const { return w_balance_; }
  void setConstraintWeights(double w_joint_limits, double w_control_limits);

  bool loadReferences(const std::string& q_ref_path, const std::string& v_ref_path);
  void getReferenceWindow(int t0, int N, std::vector<Eigen::VectorXd>& x_ref_window,
                          std::vector<Eigen::VectorXd>& u_ref_window, std::vector<Eigen::Vector3d>& com_ref_window) const;
  bool loadContactSchedule(const std::string& contact_path);
  bool isStance(int ee_idx, int t) const;
  int jointId(const std::string& name) const;
  std::string getEEFrameName(int ee_idx) const;
  Eigen::Vector3d getEEReference(int t, int ee_idx) const;
  Eigen::Vector3d getEEVelReference(int t, int ee_idx) const;
  Eigen::Vector3d getCoMVelReference(int t) const;
  void resetToReference(int t);
  void scaleRobotMass(double scale_factor);
  Eigen::Vector3d computeCoM(const Eigen::VectorXd& x) const;
  void initializeStandingPose();
  void computeGravComp(Eigen::VectorXd& ugrav) const;

  mjModel* model() const { return const_cast<mjModel*>(&model_); }
  mjData* data() const { return const_cast<mjData*>(&data_); }

  // ---- used by the iLQR / MPC shims (not part of the reference API) ----
  H1Ilqr* query_handle() const { return query_; }                 // batch-1 handle for plant / FK queries
  const H1Model& dynamics_model() const { return dyn_model_; }
  int model_version() const { return model_version_; }
  bool model_from_file() const { return model_from_file_; }       // loadModel parsed the MJCF (otherwise: built-in tables)
  const std::vector<std::string>& joint_names() const { return joint_names_; }
  const std::vector<std::string>& body_names() const { return body_names_; }            // bumped by setTimeStep / setGravity / scaleRobotMass
  H1Weights weights() const;                                      // current Q/R/Qf diagonals + task weights
  bool weights_are_diagonal() const { return diag_ok_; }          // (kept name) true when Q, R, Qf are symmetric
  bool push_weights(H1Ilqr* h) const;                             // set_weights (+ set_weight_matrices for full Q / R / Qf)
  int reference_rows() const { return static_cast<int>(x_ref_full_.size()); }
  const std::vector<std::vector<int>>& contact_schedule() const { return contact_schedule_; }
  void refresh_bias();                                            // data_.qfrc_bias <- GPU

 private:
  bool ensure_query() const;
  void model_changed();
  bool loaded_;
  int model_version_;
  bool model_from_file_ = false;
  std::vector<std::string> joint_names_, body_names_;
  int nx_, nu_;
  double dt_;
  H1Model dyn_model_;
  mutable H1Ilqr* query_;
  mjModel model_;
  mjData data_;
  std::vector<double> qpos_, qvel_, ctrl_, qfrc_bias_;
  Eigen::MatrixXd Q_, R_, Qf_;
  bool diag_ok_, full_weights_;
  double w_com_, w_com_vel_, w_ee_pos_, w_ee_vel_, w_joint_limits_, w_control_limits_, w_upright_, w_balance_;
  std::vector<Eigen::VectorXd> x_ref_full_, u_ref_full_;
  std::vector<Eigen::Vector3d> com_ref_full_, com_vel_ref_full_;
  std::vector<std::vector<Eigen::Vector3d>> ee_pos_ref_full_, ee_vel_ref_full_;
  std::vector<std::vector<int>> contact_schedule_;
};
