// iLQR shim over the C ABI (reference: src/ilqr/ilqr.cpp). All numerical stages run on the GPU.
#include "ilqr/ilqr.hpp"
#include "common/model_loader.hpp"
#include <iostream>
#include <stdexcept>

iLQR::iLQR(RobotUtils& robot, int N, double dt, const std::string& urdf_path)
    : robot_(robot), N_(N), dt_(dt), h_(nullptr) {
  // The cost model is read from the URDF at run time (reference: symDerivatives builds its Pinocchio model from urdf_path,
  // derivatives.cpp:26-39), keyed by the joint names of the MJCF the robot parsed; otherwise the built-in tables are used.
  std::string err;
  if (robot_.model_from_file() &&
      load_urdf_model(urdf_path, *h1_default_cost_model(), robot_.joint_names(), robot_.body_names(), &cost_model_, &err)) {
    cost_from_file_ = true;
    std::cout << "Parsed URDF cost model: " << urdf_path << std::endl;
  } else {
    cost_from_file_ = false;
    if (robot_.model_from_file()) std::cerr << "Note: " << urdf_path << " not usable (" << err << "); using the built-in H1 cost model" << std::endl;
  }
  h1ilqr_default_options(&opt_);
  const int nx = robot_.nx(), nu = robot_.nu();
  xbar_.assign(N_ + 1, Eigen::VectorXd::Zero(nx));
  ubar_.assign(N_, Eigen::VectorXd::Zero(nu));
  kff_.assign(N_, Eigen::VectorXd::Zero(nu));
  K_.assign(N_, Eigen::MatrixXd::Zero(nu, nx));
  if (!recreate()) throw std::runtime_error(std::string("iLQR: cannot create the GPU solver: ") + h1ilqr_last_error());
  std::cout << "iLQR initialized with horizon N=" << N_ << ", dt=" << dt_ << std::endl;
}
iLQR::~iLQR() { if (h_) h1ilqr_destroy(h_); }

// (Re)build the device handle from the robot's CURRENT dynamics model and the current options; the regularisation carries
// over. The old handle survives a failure (e.g. max_iterations beyond H1ILQR_MAX_ITERS), so the solver stays usable.
bool iLQR::recreate() {
  double lambda = opt_.reg_init;
  if (h_) h1ilqr_get_regularization(h_, &lambda);
  H1Model dm = robot_.dynamics_model();
  H1Ilqr* fresh = nullptr;
  if (h1ilqr_create(&dm, cost_from_file_ ? &cost_model_ : nullptr, &opt_, 1, N_, 0, &fresh) != H1ILQR_OK) return false;
  if (h1ilqr_set_regularization(fresh, &lambda, 1) != H1ILQR_OK) { h1ilqr_destroy(fresh); return false; }
  if (h_) h1ilqr_destroy(h_);
  h_ = fresh;
  model_version_ = robot_.model_version();
  return true;
}
// RobotUtils::setTimeStep / setGravity / scaleRobotMass after construction change the model the plant steps with; the
// solver must optimise with the same one (the reference shares one mjModel between both).
void iLQR::sync_model() {
  if (model_version_ != robot_.model_version() && !recreate())
    throw std::runtime_error(std::string("iLQR: cannot rebuild the GPU solver for the changed model: ") + h1ilqr_last_error());
}
void iLQR::setRegularization(double lambda) { if (h_) h1ilqr_set_regularization(h_, &lambda, 1); }
void iLQR::setMaxIterations(int max_iter) {
  const int old = opt_.max_iterations;
  opt_.max_iterations = max_iter;
  if (!recreate()) { opt_.max_iterations = old; throw std::runtime_error(std::string("iLQR::setMaxIterations: ") + h1ilqr_last_error()); }
}
void iLQR::setTolerance(double tol) {
  const double old = opt_.tolerance;
  opt_.tolerance = tol;
  if (!recreate()) { opt_.tolerance = old; throw std::runtime_error(std::string("iLQR::setTolerance: ") + h1ilqr_last_error()); }
}

bool iLQR::upload_window(const std::vector<Eigen::VectorXd>& x_ref, const std::vector<Eigen::VectorXd>& u_ref,
                         const std::vector<Eigen::Vector3d>& com_ref) {
  const int nx = robot_.nx(), nu = robot_.nu();
  std::vector<double> xr((N_ + 1) * nx), ur(N_ * nu), cr((N_ + 1) * 3), er((N_ + 1) * 6), cv((N_ + 1) * 3, 0.0);
  std::vector<int> st((N_ + 1) * 2);
  for (int t = 0; t <= N_; ++t) {
    for (int i = 0; i < nx; ++i) xr[t * nx + i] = x_ref[t](i);
    for (int i = 0; i < 3; ++i) cr[t * 3 + i] = com_ref[t](i);
    // horizon-local lookups, exactly as the reference does (quirk Q6); getEEReference throws past the table
    for (int e = 0; e < 2; ++e) {
      st[t * 2 + e] = robot_.isStance(e, t) ? 1 : 0;
      Eigen::Vector3d p = robot_.getEEReference(t, e);
      for (int i = 0; i < 3; ++i) er[t * 6 + e * 3 + i] = p(i);
    }
    if (robot_.getCoMVelWeight() > 0.0) { Eigen::Vector3d v = robot_.getCoMVelReference(t); for (int i = 0; i < 3; ++i) cv[t * 3 + i] = v(i); }
    if (t < N_) for (int i = 0; i < nu; ++i) ur[t * nu + i] = u_ref[t](i);
  }
  if (!robot_.push_weights(h_)) return false;
  return h1ilqr_set_reference_window(h_, xr.data(), ur.data(), cr.data(), er.data(), st.data(), cv.data(), 1) == H1ILQR_OK;
}

void iLQR::download_solution() {
  const int nx = robot_.nx(), nu = robot_.nu();
  std::vector<double> xb((N_ + 1) * nx), ub(N_ * nu), K(N_ * nu * nx), kf(N_ * nu);
  h1ilqr_get_trajectory(h_, xb.data(), ub.data());
  h1ilqr_get_gains(h_, K.data(), kf.data());
  for (int t = 0; t <= N_; ++t) for (int i = 0; i < nx; ++i) xbar_[t](i) = xb[t * nx + i];
  for (int t = 0; t < N_; ++t) {
    for (int i = 0; i < nu; ++i) { ubar_[t](i) = ub[t * nu + i]; kff_[t](i) = kf[t * nu + i]; }
    for (int j = 0; j < nx; ++j) for (int i = 0; i < nu; ++i) K_[t](i, j) = K[(t * nx + j) * nu + i];
  }
}

void iLQR::initializeWithReference(const Eigen::VectorXd& x0, const std::vector<Eigen::VectorXd>& x_ref,
                                   const std::vector<Eigen::VectorXd>& u_ref, const std::vector<Eigen::Vector3d>& com_ref,
                                   const std::vector<Eigen::VectorXd>* prev_xbar, const std::vector<Eigen::VectorXd>* prev_ubar) {
  (void)x_ref; (void)u_ref; (void)com_ref;
  sync_model();
  const int nx = robot_.nx(), nu = robot_.nu();
  if (prev_xbar && prev_ubar && prev_xbar->size() == xbar_.size() && prev_ubar->size() == ubar_.size()) {
    // warm start: the caller's previous solution goes to the device, which shifts it by one knot and rolls out the last
    // step (ilqr.cpp:68-81) — one upload and one launch sequence instead of a host-side shift
    std::vector<double> xb((N_ + 1) * nx), ub(N_ * nu);
    for (int t = 0; t <= N_; ++t) for (int i = 0; i < nx; ++i) xb[t * nx + i] = (*prev_xbar)[t](i);
    for (int t = 0; t < N_; ++t) for (int i = 0; i < nu; ++i) ub[t * nu + i] = (*prev_ubar)[t](i);
    const int warm = 1;
    if (h1ilqr_set_previous_solution(h_, xb.data(), ub.data()) != H1ILQR_OK ||
        h1ilqr_initialize(h_, x0.data(), &warm, nullptr, 1) != H1ILQR_OK)
      throw std::runtime_error(std::string("iLQR warm start: ") + h1ilqr_last_error());
  } else {
    std::cout << "Initial Guess Strategy: Gravity Compensation" << std::endl;
    Eigen::VectorXd ug;
    robot_.computeGravComp(ug);
    h1ilqr_initialize(h_, x0.data(), nullptr, ug.data(), 1);
  }
  download_solution();
}

bool iLQR::solve(const Eigen::VectorXd& x0, const std::vector<Eigen::VectorXd>& x_ref,
                 const std::vector<Eigen::VectorXd>& u_ref, const std::vector<Eigen::Vector3d>& com_ref, double& cost_out) {
  if (x_ref.size() != (size_t)(N_ + 1) || u_ref.size() != (size_t)N_ || com_ref.size() != (size_t)(N_ + 1)) {
    std::cerr << "Reference size mismatch: x_ref=" << x_ref.size() << " expected=" << N_ + 1 << ", u_ref=" << u_ref.size()
              << " expected=" << N_ << ", com_ref=" << com_ref.size() << " expected=" << N_ + 1 << std::endl;
    return false;
  }
  sync_model();
  if (!robot_.weights_are_diagonal()) throw std::runtime_error("Q, R and Qf must be symmetric");
  if (!upload_window(x_ref, u_ref, com_ref)) throw std::runtime_error(std::string("iLQR upload: ") + h1ilqr_last_error());
  int status = 0, iters = 0;
  int rc = h1ilqr_solve(h_, x0.data(), &cost_out, &iters, &status);
  if (rc == H1ILQR_ENOTFINITE) std::cout << "Warning: Non-finite gains" << std::endl;
  else if (rc != H1ILQR_OK) throw std::runtime_error(std::string("iLQR solve: ") + h1ilqr_last_error());
  download_solution();
  return true;
}

H1StageTimes iLQR::lastStageTimes() const {
  H1StageTimes t;
  h1ilqr_get_stage_times(h_, &t);
  return t;
}
