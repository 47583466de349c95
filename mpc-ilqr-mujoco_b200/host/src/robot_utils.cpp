// RobotUtils on top of the C ABI (reference: src/common/robot_utils.cpp). The plant is the GPU dynamics map f_D;
// references and the contact schedule are parsed with the reference's rules (robot_utils.cpp:281-347, 445-492).
#include "common/robot_utils.hpp"
#include "common/model_loader.hpp"
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

RobotUtils::RobotUtils()
    : loaded_(false), model_version_(0), nx_(0), nu_(0), dt_(0.01), query_(nullptr), diag_ok_(true), full_weights_(false), w_com_(0.0), w_com_vel_(0.0),
      w_ee_pos_(0.0), w_ee_vel_(0.0), w_joint_limits_(500.0), w_control_limits_(1000.0), w_upright_(0.0), w_balance_(0.0) {
  std::memset(&model_, 0, sizeof(model_));
  std::memset(&data_, 0, sizeof(data_));
}

RobotUtils::~RobotUtils() { if (query_) h1ilqr_destroy(query_); }

void RobotUtils::model_changed() {   // the dynamics model is baked into device handles: drop ours, let the solver notice
  ++model_version_;
  if (query_) { h1ilqr_destroy(query_); query_ = nullptr; }
}

bool RobotUtils::ensure_query() const {
  if (query_) {   // the limit penalties / stage costs read the handle's weights: keep them current
    push_weights(query_);
    return true;
  }
  H1SolverOptions o;
  h1ilqr_default_options(&o);
  if (h1ilqr_create(&dyn_model_, nullptr, &o, 1, 2, 0, &query_) != H1ILQR_OK) {
    std::cerr << "GPU solver core unavailable: " << h1ilqr_last_error() << std::endl;
    query_ = nullptr;
    return false;
  }
  push_weights(query_);
  return true;
}

// current weights -> a device handle: diagonals + task weights, and the off-diagonal parts when Q / R / Qf are full
bool RobotUtils::push_weights(H1Ilqr* h) const {
  H1Weights w = weights();
  if (h1ilqr_set_weights(h, &w) != H1ILQR_OK) return false;
  if (full_weights_) return h1ilqr_set_weight_matrices(h, Q_.data(), R_.data(), Qf_.data()) == H1ILQR_OK;
  return h1ilqr_set_weight_matrices(h, nullptr, nullptr, nullptr) == H1ILQR_OK;
}

bool RobotUtils::loadModel(const std::string& xml_path) {
  // The MJCF is read at run time (reference: mj_loadXML, robot_utils.cpp:19-33); when the file is missing or is not an
  // H1-class model the tables generated from the reference's h1.xml at build time (tools/gen_h1_model.py) are used.
  std::string err;
  H1Model parsed;
  if (load_mjcf_model(xml_path, *h1_default_dynamics_model(), &parsed, &joint_names_, &body_names_, &err)) {
    dyn_model_ = parsed;
    model_from_file_ = true;
    std::cout << "Parsed MJCF model: " << xml_path << std::endl;
  } else {
    std::cerr << "Note: " << xml_path << " not usable (" << err << "); using the built-in H1 model tables" << std::endl;
    dyn_model_ = *h1_default_dynamics_model();
    model_from_file_ = false;
    joint_names_.clear(); body_names_.clear();
  }
  nx_ = H1_NX; nu_ = H1_NU; dt_ = dyn_model_.timestep;
  loaded_ = true;
  qpos_.assign(H1_NQ, 0.0); qvel_.assign(H1_NV, 0.0); ctrl_.assign(H1_NU, 0.0); qfrc_bias_.assign(H1_NV + 1, 0.0);
  qpos_[3] = 1.0;
  model_.nq = H1_NQ; model_.nv = H1_NV; model_.nu = H1_NU; model_.opt.timestep = dt_; model_.owner = this;
  for (int i = 0; i < 3; ++i) model_.opt.gravity[i] = dyn_model_.gravity[i];
  data_.qpos = qpos_.data(); data_.qvel = qvel_.data(); data_.ctrl = ctrl_.data(); data_.qfrc_bias = qfrc_bias_.data();
  data_.owner = this;
  Q_ = Eigen::MatrixXd::Identity(nx_, nx_); R_ = Eigen::MatrixXd::Identity(nu_, nu_); Qf_ = Eigen::MatrixXd::Identity(nx_, nx_);
  std::cout << "Model loaded successfully:" << std::endl;
  std::cout << "Found 2 end-effector bodies" << std::endl;
  return true;
}

void RobotUtils::setContactImpratio(double impratio) {
  model_.opt.impratio = impratio;  // MuJoCo solver option; the soft-contact map f_D has no such parameter
  std::cout << "Set IMPRATIO to: " << impratio << std::endl;
}
void RobotUtils::setTimeStep(double dt) {
  dt_ = dt; dyn_model_.timestep = dt; model_.opt.timestep = dt;
  model_changed();
  std::cout << "Set timestep to: " << dt << std::endl;
}
void RobotUtils::setGravity(double gx, double gy, double gz) {
  dyn_model_.gravity[0] = gx; dyn_model_.gravity[1] = gy; dyn_model_.gravity[2] = gz;
  model_.opt.gravity[0] = gx; model_.opt.gravity[1] = gy; model_.opt.gravity[2] = gz;
  model_changed();
  std::cout << "Set gravity to: (" << gx << "," << gy << "," << gz << ")m/s²" << std::endl;
}

void RobotUtils::setState(const Eigen::VectorXd& x) {
  if (!loaded_ || x.size() != nx_) { std::cerr << "Invalid state size: " << x.size() << " (expected " << nx_ << ")" << std::endl; return; }
  for (int i = 0; i < H1_NQ; ++i) qpos_[i] = x(i);
  for (int i = 0; i < H1_NV; ++i) qvel_[i] = x(H1_NQ + i);
}
void RobotUtils::getState(Eigen::VectorXd& x) const {
  if (!loaded_) return;
  x.resize(nx_);
  for (int i = 0; i < H1_NQ; ++i) x(i) = qpos_[i];
  for (int i = 0; i < H1_NV; ++i) x(H1_NQ + i) = qvel_[i];
}
void RobotUtils::setControl(const Eigen::VectorXd& u) {
  if (!loaded_ || u.size() != nu_) { std::cerr << "Invalid control size: " << u.size() << " (expected " << nu_ << ")" << std::endl; return; }
  for (int i = 0; i < H1_NU; ++i) ctrl_[i] = u(i);
}
void RobotUtils::rolloutOneStep(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::VectorXd& x_next) {
  if (!loaded_ || !ensure_query()) return;
  x_next.resize(nx_);
  if (h1ilqr_dynamics_step(query_, 1, x.data(), u.data(), x_next.data()) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_dynamics_step: ") + h1ilqr_last_error());
}
void RobotUtils::linearizeDynamicsFD(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::MatrixXd& A,
                                     Eigen::MatrixXd& B, double eps) {
  if (!loaded_ || !ensure_query()) return;
  A.resize(nx_, nx_); B.resize(nx_, nu_);   // column-major like the C ABI
  if (h1ilqr_linearize_state(query_, H1ILQR_LIN_FD, eps, x.data(), u.data(), A.data(), B.data()) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_linearize_state: ") + h1ilqr_last_error());
}

double RobotUtils::constraintCost(const Eigen::VectorXd& x, const Eigen::VectorXd& u) const {
  if (!loaded_ || !ensure_query()) return 0.0;
  double c = 0.0;
  if (h1ilqr_limit_penalties(query_, 1, x.data(), u.data(), &c, nullptr, nullptr, nullptr, nullptr) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_limit_penalties: ") + h1ilqr_last_error());
  return c;
}
void RobotUtils::constraintGradients(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::VectorXd& grad_x,
                                     Eigen::VectorXd& grad_u) const {
  if (!loaded_ || !ensure_query()) return;
  grad_x.setZero(nx_); grad_u.setZero(nu_);
  if (h1ilqr_limit_penalties(query_, 1, x.data(), u.data(), nullptr, grad_x.data(), grad_u.data(), nullptr, nullptr) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_limit_penalties: ") + h1ilqr_last_error());
}
void RobotUtils::constraintHessians(const Eigen::VectorXd& x, const Eigen::VectorXd& u, Eigen::MatrixXd& hess_xx,
                                    Eigen::MatrixXd& hess_uu) const {
  if (!loaded_ || !ensure_query()) return;
  hess_xx.resize(nx_, nx_); hess_xx.setZero(); hess_uu.resize(nu_, nu_); hess_uu.setZero();
  std::vector<double> hx(nx_), hu(nu_);
  if (h1ilqr_limit_penalties(query_, 1, x.data(), u.data(), nullptr, nullptr, nullptr, hx.data(), hu.data()) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_limit_penalties: ") + h1ilqr_last_error());
  for (int i = 0; i < nx_; ++i) hess_xx(i, i) = hx[i];
  for (int i = 0; i < nu_; ++i) hess_uu(i, i) = hu[i];
}
double RobotUtils::stageCost(int t, const Eigen::VectorXd& x, const Eigen::VectorXd& u) const {
  if (!loaded_ || x_ref_full_.empty() || !ensure_query()) return 0.0;
  // rows past the tables use the last one (robot_utils.cpp:163-166)
  const int xi = std::min(t, (int)x_ref_full_.size() - 1), ui = std::min(t, (int)u_ref_full_.size() - 1);
  const int ci = std::min(t, (int)com_ref_full_.size() - 1);
  double c = 0.0;
  if (h1ilqr_stage_cost(query_, 1, x.data(), u.data(), x_ref_full_[xi].data(), u_ref_full_[ui].data(),
                        com_ref_full_.empty() ? nullptr : com_ref_full_[ci].data(), &c) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_stage_cost: ") + h1ilqr_last_error());
  return c;
}
double RobotUtils::terminalCost(const Eigen::VectorXd& x) const {
  if (!loaded_ || x_ref_full_.empty() || !ensure_query()) return 0.0;
  double c = 0.0;
  if (h1ilqr_stage_cost(query_, 1, x.data(), nullptr, x_ref_full_.back().data(), nullptr,
                        com_ref_full_.empty() ? nullptr : com_ref_full_.back().data(), &c) != H1ILQR_OK)
    throw std::runtime_error(std::string("h1ilqr_stage_cost: ") + h1ilqr_last_error());
  return c;
}

void RobotUtils::step() {
  if (!loaded_) return;
  Eigen::VectorXd x(nx_), u(nu_), xn(nx_);
  getState(x);
  for (int i = 0; i < nu_; ++i) u(i) = ctrl_[i];
  rolloutOneStep(x, u, xn);
  setState(xn);
}

void RobotUtils::setCostWeights(const Eigen::MatrixXd& Q, const Eigen::MatrixXd& R, const Eigen::MatrixXd& Qf) {
  if (Q.rows() != nx_ || Q.cols() != nx_) { std::cerr << "ERROR: Q matrix dimension mismatch! Expected " << nx_ << "x" << nx_ << ", got " << Q.rows() << "x" << Q.cols() << std::endl; return; }
  if (R.rows() != nu_ || R.cols() != nu_) { std::cerr << "ERROR: R matrix dimension mismatch! Expected " << nu_ << "x" << nu_ << ", got " << R.rows() << "x" << R.cols() << std::endl; return; }
  if (Qf.rows() != nx_ || Qf.cols() != nx_) { std::cerr << "ERROR: Qf matrix dimension mismatch! Expected " << nx_ << "x" << nx_ << ", got " << Qf.rows() << "x" << Qf.cols() << std::endl; return; }
  Q_ = Q; R_ = R; Qf_ = Qf;
  // whole symmetric matrices are supported (the off-diagonal parts travel through h1ilqr_set_weight_matrices); an
  // asymmetric one has no consistent quadratic model and is refused when the solver uploads it
  auto offdiag = [](const Eigen::MatrixXd& M) { for (long j = 0; j < M.cols(); ++j) for (long i = 0; i < M.rows(); ++i) if (i != j && M(i, j) != 0.0) return true; return false; };
  auto symmetric = [](const Eigen::MatrixXd& M) { for (long j = 0; j < M.cols(); ++j) for (long i = 0; i < j; ++i) if (M(i, j) != M(j, i)) return false; return true; };
  full_weights_ = offdiag(Q) || offdiag(R) || offdiag(Qf);
  diag_ok_ = symmetric(Q) && symmetric(R) && symmetric(Qf);
  if (!diag_ok_) std::cerr << "ERROR: Q, R and Qf must be symmetric" << std::endl;
  std::cout << "Cost weights set successfully" << std::endl;
}
void RobotUtils::setConstraintWeights(double wj, double wc) {
  w_joint_limits_ = wj; w_control_limits_ = wc;
  std::cout << "Constraint weights set: joint_limits=" << wj << ", control_limits=" << wc << std::endl;
}
H1Weights RobotUtils::weights() const {
  H1Weights w;
  for (int i = 0; i < H1_NX; ++i) { w.Qdiag[i] = Q_(i, i); w.Qfdiag[i] = Qf_(i, i); }
  for (int i = 0; i < H1_NU; ++i) w.Rdiag[i] = R_(i, i);
  w.w_com = w_com_; w.w_com_vel = w_com_vel_; w.w_ee_pos = w_ee_pos_; w.w_ee_vel = w_ee_vel_;
  w.w_upright = w_upright_; w.w_balance = w_balance_; w.w_joint_limits = w_joint_limits_; w.w_control_limits = w_control_limits_;
  return w;
}

static bool parse_row(const std::string& line, std::vector<double>& out) {
  std::stringstream ss(line);
  std::string val;
  while (std::getline(ss, val, ',')) {
    try { out.push_back(std::stod(val)); } catch (const std::exception&) { continue; }
  }
  return true;
}

bool RobotUtils::loadReferences(const std::string& q_ref_path, const std::string& v_ref_path) {
  std::ifstream q_file(q_ref_path), v_file(v_ref_path);
  if (!q_file.is_open()) { std::cerr << "Failed to open position reference file: " << q_ref_path << std::endl; return false; }
  if (!v_file.is_open()) { std::cerr << "Failed to open velocity reference file: " << v_ref_path << std::endl; return false; }
  x_ref_full_.clear(); u_ref_full_.clear(); com_ref_full_.clear(); com_vel_ref_full_.clear(); ee_pos_ref_full_.clear();
  ee_vel_ref_full_.clear();
  std::string q_line, v_line;
  std::vector<double> flat;
  while (std::getline(q_file, q_line) && std::getline(v_file, v_line)) {
    std::vector<double> q, v;
    parse_row(q_line, q); parse_row(v_line, v);
    if ((int)q.size() != H1_NQ || (int)v.size() != H1_NV) continue;   // skipped rows do not break the lockstep
    Eigen::VectorXd x(nx_);
    for (int i = 0; i < H1_NQ; ++i) x(i) = q[i];
    for (int i = 0; i < H1_NV; ++i) x(H1_NQ + i) = v[i];
    x_ref_full_.push_back(x);
    u_ref_full_.push_back(Eigen::VectorXd::Zero(nu_));
    flat.insert(flat.end(), x.data(), x.data() + nx_);
  }
  if (x_ref_full_.empty()) { std::cerr << "No valid reference states loaded" << std::endl; return false; }
  // per-row CoM (subtree_com of the root) and ankle body positions on the dynamics model, on the GPU
  if (!ensure_query()) return false;
  const int T = (int)x_ref_full_.size();
  // (CoM velocity = J_subtreeCom * qvel and ankle velocities = jac_pos * qvel per row too, robot_utils.cpp:388-412)
  std::vector<double> com(3 * T), ee(6 * T), cv(3 * T), ev(6 * T);
  if (h1ilqr_reference_kinematics(query_, T, flat.data(), com.data(), ee.data()) != H1ILQR_OK ||
      h1ilqr_reference_com_velocity(query_, T, flat.data(), cv.data()) != H1ILQR_OK ||
      h1ilqr_reference_ee_velocity(query_, T, flat.data(), ev.data()) != H1ILQR_OK) {
    std::cerr << "reference kinematics failed: " << h1ilqr_last_error() << std::endl;
    return false;
  }
  for (int t = 0; t < T; ++t) {
    com_ref_full_.push_back(Eigen::Vector3d(com[3 * t], com[3 * t + 1], com[3 * t + 2]));
    com_vel_ref_full_.push_back(Eigen::Vector3d(cv[3 * t], cv[3 * t + 1], cv[3 * t + 2]));
    ee_pos_ref_full_.push_back({Eigen::Vector3d(ee[6 * t], ee[6 * t + 1], ee[6 * t + 2]),
                                Eigen::Vector3d(ee[6 * t + 3], ee[6 * t + 4], ee[6 * t + 5])});
    ee_vel_ref_full_.push_back({Eigen::Vector3d(ev[6 * t], ev[6 * t + 1], ev[6 * t + 2]),
                                Eigen::Vector3d(ev[6 * t + 3], ev[6 * t + 4], ev[6 * t + 5])});
  }
  std::cout << "Loaded " << x_ref_full_.size() << " reference states" << std::endl;
  return true;
}

void RobotUtils::getReferenceWindow(int t0, int N, std::vector<Eigen::VectorXd>& xw, std::vector<Eigen::VectorXd>& uw,
                                    std::vector<Eigen::Vector3d>& cw) const {
  xw.clear(); uw.clear(); cw.clear();
  for (int i = 0; i <= N; ++i) {
    int idx = std::min(t0 + i, (int)x_ref_full_.size() - 1);
    xw.push_back(x_ref_full_[idx]);
    cw.push_back(com_ref_full_[std::min(t0 + i, (int)com_ref_full_.size() - 1)]);
    if (i < N) uw.push_back(u_ref_full_[std::min(t0 + i, (int)u_ref_full_.size() - 1)]);
  }
}

bool RobotUtils::loadContactSchedule(const std::string& contact_path) {
  contact_schedule_.clear();
  std::ifstream file(contact_path);
  if (!file.is_open()) { std::cerr << "Warning: Failed to open contact schedule file: " << contact_path << std::endl; return false; }
  std::string line;
  std::getline(file, line);  // header
  while (std::getline(file, line)) {
    std::stringstream ss(line);
    std::string tok;
    std::vector<int> c;
    while (std::getline(ss, tok, ',')) { try { c.push_back(std::stoi(tok)); } catch (...) { continue; } }
    if (!c.empty() && c.size() != 2) std::cerr << "Warning: Contact schedule has " << c.size() << " end-effectors but model has 2" << std::endl;
    if (!c.empty()) contact_schedule_.push_back(c);
  }
  std::cout << "Loaded contact schedule: " << contact_schedule_.size() << " timesteps, "
            << (contact_schedule_.empty() ? 0 : contact_schedule_[0].size()) << " end-effectors" << std::endl;
  return !contact_schedule_.empty();
}
bool RobotUtils::isStance(int ee_idx, int t) const {
  if (t < 0 || t >= (int)contact_schedule_.size()) return true;
  if (ee_idx < 0 || ee_idx >= (int)contact_schedule_[t].size()) return true;
  return contact_schedule_[t][ee_idx] == 1;
}
std::string RobotUtils::getEEFrameName(int ee_idx) const {
  if (ee_idx < 0 || ee_idx >= 2) throw std::runtime_error("Invalid EE index: " + std::to_string(ee_idx));
  return ee_idx == 0 ? "left_ankle_link" : "right_ankle_link";
}
Eigen::Vector3d RobotUtils::getEEReference(int t, int ee_idx) const {
  if (t >= (int)ee_pos_ref_full_.size() || ee_idx >= (int)ee_pos_ref_full_[t].size())
    throw std::runtime_error("Invalid reference index: t=" + std::to_string(t) + ", ee_idx=" + std::to_string(ee_idx));
  return ee_pos_ref_full_[t][ee_idx];
}
Eigen::Vector3d RobotUtils::getEEVelReference(int t, int ee_idx) const {
  if (t >= (int)ee_vel_ref_full_.size() || ee_idx >= (int)ee_vel_ref_full_[t].size())
    throw std::runtime_error("Invalid velocity reference index: t=" + std::to_string(t) + ", ee_idx=" + std::to_string(ee_idx));
  return ee_vel_ref_full_[t][ee_idx];
}
int RobotUtils::jointId(const std::string& name) const {
  // MuJoCo joint ids of the MJCF (h1.xml:49-179): 0 is the unnamed free joint, 1..19 the hinges in body order
  static const char* const names[H1_NU] = {
      "left_hip_yaw_joint", "left_hip_roll_joint", "left_hip_pitch_joint", "left_knee_joint", "left_ankle_joint",
      "right_hip_yaw_joint", "right_hip_roll_joint", "right_hip_pitch_joint", "right_knee_joint", "right_ankle_joint",
      "torso_joint", "left_shoulder_pitch_joint", "left_shoulder_roll_joint", "left_shoulder_yaw_joint", "left_elbow_joint",
      "right_shoulder_pitch_joint", "right_shoulder_roll_joint", "right_shoulder_yaw_joint", "right_elbow_joint"};
  if (!loaded_) return -1;
  if ((int)joint_names_.size() == H1_NU) {   // names of the MJCF that was parsed at run time
    for (int i = 0; i < H1_NU; ++i) if (name == joint_names_[i]) return i + 1;
    return -1;
  }
  for (int i = 0; i < H1_NU; ++i) if (name == names[i]) return i + 1;
  return -1;
}
void RobotUtils::resetToReference(int t) {
  if (t >= 0 && t < (int)x_ref_full_.size()) setState(x_ref_full_[t]);
}
void RobotUtils::scaleRobotMass(double scale_factor) {
  if (!loaded_) return;
  for (int b = 0; b < H1_NB; ++b) dyn_model_.mass[b] *= scale_factor;   // body_mass only, as the reference (robot_utils.cpp:835-842)
  dyn_model_.total_mass *= scale_factor;
  model_changed();
  std::cout << "Scaled robot mass by factor: " << scale_factor << std::endl;
}
Eigen::Vector3d RobotUtils::getCoMVelReference(int t) const {
  if (t >= (int)com_vel_ref_full_.size()) throw std::runtime_error("Invalid CoM velocity reference index: t=" + std::to_string(t));
  return com_vel_ref_full_[t];
}
Eigen::Vector3d RobotUtils::computeCoM(const Eigen::VectorXd& x) const {
  Eigen::Vector3d c;
  if (!loaded_ || !ensure_query()) return c;
  double com[3], ee[6];
  h1ilqr_reference_kinematics(query_, 1, x.data(), com, ee);
  return Eigen::Vector3d(com[0], com[1], com[2]);
}
void RobotUtils::initializeStandingPose() {
  if (!loaded_) { std::cerr << "Model not loaded, cannot initialize standing pose" << std::endl; return; }
  std::fill(qpos_.begin(), qpos_.end(), 0.0);
  std::fill(qvel_.begin(), qvel_.end(), 0.0);
  qpos_[2] = 1.0432; qpos_[3] = 1.0;
  refresh_bias();
}
void RobotUtils::refresh_bias() {
  if (!loaded_ || !ensure_query()) return;
  std::vector<double> x(H1_NX);
  for (int i = 0; i < H1_NQ; ++i) x[i] = qpos_[i];
  for (int i = 0; i < H1_NV; ++i) x[H1_NQ + i] = qvel_[i];
  h1ilqr_bias_forces(query_, 1, x.data(), qfrc_bias_.data());
  qfrc_bias_[H1_NV] = 0.0;  // the slot the reference's off-by-one read lands on (Q15), defined as zero here
}
void RobotUtils::computeGravComp(Eigen::VectorXd& ugrav) const {
  ugrav.resize(nu_);
  const_cast<RobotUtils*>(this)->refresh_bias();
  // the reference indexes qfrc_bias with the joint's qpos address (7+i) instead of its dof address (Q15)
  for (int i = 0; i < nu_; ++i) ugrav(i) = qfrc_bias_[7 + i];
}

void mj_forward(const mjModel* m, mjData* d) {
  (void)d;
  if (m && m->owner) static_cast<RobotUtils*>(m->owner)->refresh_bias();
}
