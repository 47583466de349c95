// TEST INFRASTRUCTURE ONLY — CPU oracle for the H1 iLQR hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
// Parity status: UNPINNED (the reference ships no tests or golden vectors, and its third-party
// dependencies MuJoCo / Pinocchio / CasADi / Eigen are absent here). See oracle/README.md.
#pragma once
#include "../include/h1ilqr.h"
#include <vector>

namespace orc {

// ---- dynamics (oracle_dynamics.cpp) ----
void normalize_quat(const double* q, double* qn);
void dyn_fk(const H1Model& md, const double* q, double (*R)[9], double (*r)[3]);
void dyn_com(const H1Model& md, const double* x, double* com);
void dyn_body_pos(const H1Model& md, const double* x, int body, double* p);
void dyn_com_vel(const H1Model& md, const double* x, double* com_vel);
void dyn_bias(const H1Model& md, const double* x, double* bias);
void dyn_step(const H1Model& md, const double* x, const double* u, double* xn);
void dyn_linearize_fd(const H1Model& md, const double* x, const double* u, double eps, double* A, double* B);
void dyn_linearize_ad(const H1Model& md, const double* x, const double* u, double* A, double* B);

// ---- cost terms (oracle_cost.cpp) ----
// All gradients/Hessians are w.r.t. the Pinocchio-ordered state x~ (quaternion x,y,z,w at 3..6) and are
// ADDED to g[51] / H[51*51] (column-major) at the same indices — reference quirk Q3.
enum CostTerm { TERM_COM = 0, TERM_COM_VEL = 1, TERM_EE_POS = 2, TERM_EE_VEL = 3, TERM_UPRIGHT = 4, TERM_BALANCE = 5 };
// value of a single term (used for FD self-checks); target has 3 (2 for balance, unused for upright) entries
double cost_term_value(const H1Model& cm, int term, int ee, const double* x_mj, const double* target, double w);
// exact derivatives by second-order forward AD through a Pinocchio-style local-frame FK (slow, faithful)
void cost_term_ad(const H1Model& cm, int term, int ee, const double* x_mj, const double* target, double w,
                  double* g, double* H);
// hand-derived exact derivatives (fast); must agree with cost_term_ad to rounding
void cost_term_analytic(const H1Model& cm, int term, int ee, const double* x_mj, const double* target, double w,
                        double* g, double* H);

double limit_cost(const H1Model& md, const H1Weights& w, const double* x, const double* u);
void limit_derivs(const H1Model& md, const H1Weights& w, const double* x, const double* u, double* gx, double* gu,
                  double* hxx_diag, double* huu_diag);

// ---- problem / solver (oracle_ilqr.cpp) ----
struct Problem {
  int N = 0;
  H1Model dyn, cost;
  H1Weights w;
  H1SolverOptions opt;
  std::vector<double> x_ref, u_ref, com_ref, ee_ref, com_vel_ref;  // [(N+1)*51], [N*19], [(N+1)*3], [(N+1)*6], [(N+1)*3]
  std::vector<int> stance;                                        // [(N+1)*2]
  bool use_ad = false;  // cost derivatives through the AD path instead of the analytic path
  // full symmetric Q, R, Qf (column-major; empty = the diagonal matrices of `w`): iLQR multiplies whole matrices
  // (/root/reference/src/ilqr/ilqr.cpp:145-150, 372-373, 441)
  std::vector<double> Qfull, Rfull, Qffull;
};

struct Solver {
  Problem* p = nullptr;
  int N = 0;
  double lambda = 1e-6;
  std::vector<double> xbar, ubar, K, kff, A, B, lx, lu, lxx, luu;
  std::vector<double> cost_trace;
  std::vector<int> alpha_trace;  // [iter][2]: alpha index of first / second line search (-1 none, -2 not run)
  // decision margins of the last solve, in cost units (test diagnostics: how far every accept / reject / stop
  // decision was from flipping). ls_margin [iter][2]: min over the candidates evaluated by that line search of
  // |c - (baseline - accept_margin)| (-1 = not run); stop_margin [iter]: distance of the iteration's stop tests
  // (|cur - prev| vs tolerance, cur vs divergence_cost) from flipping (-1 = not evaluated).
  std::vector<double> ls_margin, stop_margin;
  double last_ls_margin = -1.0;
  int iters = 0;
  // MPC state
  bool has_prev = false;
  std::vector<double> prev_xbar, prev_ubar;
  void init(Problem* prob);
};

void rollout_nominal(Solver& s, const double* x0);
void linearize(Solver& s);
void cost_quadratics(Solver& s);
void backward_pass(Solver& s);
bool line_search(Solver& s, const double* x0, double* new_cost, int* alpha_index);
double total_cost(const Solver& s, const double* xtraj, const double* utraj);
void initialize(Solver& s, const double* x0, bool warm, const double* u_init);
bool solve(Solver& s, const double* x0, double* cost_out);
bool mpc_step(Solver& s, const double* x_meas, const double* u_init, double* u_apply, double* cost_out);

}  // namespace orc
