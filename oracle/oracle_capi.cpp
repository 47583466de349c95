// TEST INFRASTRUCTURE ONLY — extern "C" surface of the CPU oracle for ctypes (tests/, bench.py cpu_baseline).
#include "oracle.hpp"
#include "../include/h1_model_data.h"
#include <chrono>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

using namespace orc;

namespace {
struct Handle {
  Problem prob;
  std::vector<Solver> solvers;  // one per instance
};
const H1Model* pick(const H1Model* m, const H1Model& dflt) { return m ? m : &dflt; }
}  // namespace

extern "C" {

const H1Model* orc_default_dynamics_model(void) { return &H1_DYNAMICS_MODEL; }
const H1Model* orc_default_cost_model(void) { return &H1_COST_MODEL; }

void orc_default_options(H1SolverOptions* o) {
  o->max_iterations = 10; o->tolerance = 1e-4; o->reg_init = 1e-6; o->reg_min = 1e-6; o->reg_max = 1e-3;
  o->accept_margin = 1e-6; o->fd_eps = 1e-5; o->divergence_cost = 1e6; o->linearization = H1ILQR_LIN_ANALYTIC;
  const double a[H1ILQR_NALPHA] = {1.0, 0.8, 0.6, 0.4, 0.2, 0.1, 0.05, 0.01};
  std::memcpy(o->alphas, a, sizeof(a));
}

void orc_dyn_step(const H1Model* m, int n, const double* x, const double* u, double* xn) {
  const H1Model* md = pick(m, H1_DYNAMICS_MODEL);
  for (int i = 0; i < n; ++i) dyn_step(*md, x + i * H1_NX, u + i * H1_NU, xn + i * H1_NX);
}
void orc_dyn_linearize(const H1Model* m, const double* x, const double* u, double eps, double* A, double* B) {
  dyn_linearize_fd(*pick(m, H1_DYNAMICS_MODEL), x, u, eps, A, B);
}
void orc_dyn_linearize_ad(const H1Model* m, const double* x, const double* u, double* A, double* B) {
  dyn_linearize_ad(*pick(m, H1_DYNAMICS_MODEL), x, u, A, B);
}
void orc_dyn_com(const H1Model* m, const double* x, double* com) { dyn_com(*pick(m, H1_DYNAMICS_MODEL), x, com); }
void orc_dyn_com_vel(const H1Model* m, const double* x, double* cv) { dyn_com_vel(*pick(m, H1_DYNAMICS_MODEL), x, cv); }
void orc_dyn_bias(const H1Model* m, const double* x, double* bias) { dyn_bias(*pick(m, H1_DYNAMICS_MODEL), x, bias); }
void orc_dyn_body_pos(const H1Model* m, const double* x, int body, double* p) {
  dyn_body_pos(*pick(m, H1_DYNAMICS_MODEL), x, body, p);
}
// mode 0: value only; 1: AD derivatives; 2: analytic derivatives. g[51], H[51*51] are ACCUMULATED into.
double orc_cost_term(const H1Model* m, int term, int ee, const double* x, const double* target, double w, int mode,
                     double* g, double* H) {
  const H1Model* cm = pick(m, H1_COST_MODEL);
  if (mode == 1) cost_term_ad(*cm, term, ee, x, target, w, g, H);
  if (mode == 2) cost_term_analytic(*cm, term, ee, x, target, w, g, H);
  return cost_term_value(*cm, term, ee, x, target, w);
}
double orc_limit_cost(const H1Model* m, const H1Weights* w, const double* x, const double* u) {
  return limit_cost(*pick(m, H1_DYNAMICS_MODEL), *w, x, u);
}

void* orc_create(const H1Model* dyn, const H1Model* cost, const H1Weights* w, const H1SolverOptions* opt, int batch, int N) {
  Handle* h = new Handle;
  h->prob.N = N;
  h->prob.dyn = *pick(dyn, H1_DYNAMICS_MODEL);
  h->prob.cost = *pick(cost, H1_COST_MODEL);
  h->prob.w = *w;
  if (opt) h->prob.opt = *opt; else orc_default_options(&h->prob.opt);
  h->prob.x_ref.assign((N + 1) * H1_NX, 0.0); h->prob.u_ref.assign(N * H1_NU, 0.0);
  h->prob.com_ref.assign((N + 1) * 3, 0.0); h->prob.ee_ref.assign((N + 1) * 6, 0.0);
  h->prob.com_vel_ref.assign((N + 1) * 3, 0.0); h->prob.stance.assign((N + 1) * 2, 1);
  h->solvers.resize(batch);
  for (auto& s : h->solvers) s.init(&h->prob);
  return h;
}
// whole cost matrices (column-major); NULL = back to the diagonals of the weights struct
void orc_set_weight_matrices(void* hv, const double* Q, const double* R, const double* Qf) {
  Problem& p = static_cast<Handle*>(hv)->prob;
  p.Qfull.clear(); p.Rfull.clear(); p.Qffull.clear();
  if (!Q || !R || !Qf) return;
  p.Qfull.assign(Q, Q + H1_NX * H1_NX); p.Rfull.assign(R, R + H1_NU * H1_NU); p.Qffull.assign(Qf, Qf + H1_NX * H1_NX);
}
void orc_destroy(void* hv) { delete static_cast<Handle*>(hv); }
void orc_use_ad(void* hv, int use) { static_cast<Handle*>(hv)->prob.use_ad = use != 0; }

// one reference window shared by all instances of this handle
void orc_set_reference_window(void* hv, const double* x_ref, const double* u_ref, const double* com_ref,
                              const double* ee_ref, const int* stance, const double* com_vel_ref) {
  Handle* h = static_cast<Handle*>(hv);
  int N = h->prob.N;
  std::memcpy(h->prob.x_ref.data(), x_ref, sizeof(double) * (N + 1) * H1_NX);
  std::memcpy(h->prob.u_ref.data(), u_ref, sizeof(double) * N * H1_NU);
  std::memcpy(h->prob.com_ref.data(), com_ref, sizeof(double) * (N + 1) * 3);
  std::memcpy(h->prob.ee_ref.data(), ee_ref, sizeof(double) * (N + 1) * 6);
  std::memcpy(h->prob.stance.data(), stance, sizeof(int) * (N + 1) * 2);
  if (com_vel_ref) std::memcpy(h->prob.com_vel_ref.data(), com_vel_ref, sizeof(double) * (N + 1) * 3);
}

static Solver& S(void* hv, int i) { return static_cast<Handle*>(hv)->solvers[i]; }

void orc_initialize(void* hv, int i, const double* x0, int warm, const double* u_init) { initialize(S(hv, i), x0, warm != 0, u_init); }
void orc_rollout_nominal(void* hv, int i, const double* x0) { rollout_nominal(S(hv, i), x0); }
void orc_linearize(void* hv, int i) { linearize(S(hv, i)); }
void orc_cost_quadratics(void* hv, int i) { cost_quadratics(S(hv, i)); }
void orc_backward_pass(void* hv, int i) { backward_pass(S(hv, i)); }
int orc_line_search(void* hv, int i, const double* x0, double* new_cost, int* alpha_index) {
  return line_search(S(hv, i), x0, new_cost, alpha_index) ? 1 : 0;
}
double orc_total_cost(void* hv, int i) { Solver& s = S(hv, i); return total_cost(s, s.xbar.data(), s.ubar.data()); }
int orc_solve(void* hv, int i, const double* x0, double* cost_out) { return solve(S(hv, i), x0, cost_out) ? 1 : 0; }
int orc_mpc_step(void* hv, int i, const double* x, const double* u_init, double* u_apply, double* cost_out) {
  return mpc_step(S(hv, i), x, u_init, u_apply, cost_out) ? 1 : 0;
}
void orc_mpc_reset(void* hv, int i) { S(hv, i).has_prev = false; S(hv, i).lambda = static_cast<Handle*>(hv)->prob.opt.reg_init; }
int orc_iters(void* hv, int i) { return S(hv, i).iters; }
double orc_get_lambda(void* hv, int i) { return S(hv, i).lambda; }
void orc_set_lambda(void* hv, int i, double l) { S(hv, i).lambda = l; }

#define COPY_OUT(name, vec) \
  void orc_get_##name(void* hv, int i, double* out) { auto& v = S(hv, i).vec; std::memcpy(out, v.data(), sizeof(double) * v.size()); } \
  void orc_set_##name(void* hv, int i, const double* in) { auto& v = S(hv, i).vec; std::memcpy(v.data(), in, sizeof(double) * v.size()); }
COPY_OUT(xbar, xbar) COPY_OUT(ubar, ubar) COPY_OUT(K, K) COPY_OUT(kff, kff) COPY_OUT(A, A) COPY_OUT(B, B)
COPY_OUT(lx, lx) COPY_OUT(lu, lu) COPY_OUT(lxx, lxx) COPY_OUT(luu, luu)

void orc_get_trace(void* hv, int i, double* cost_trace, int* alpha_trace) {
  Solver& s = S(hv, i);
  std::memcpy(cost_trace, s.cost_trace.data(), sizeof(double) * s.cost_trace.size());
  std::memcpy(alpha_trace, s.alpha_trace.data(), sizeof(int) * s.alpha_trace.size());
}

// decision margins of the last solve (test diagnostics): ls_margin [max_iterations][2], stop_margin [max_iterations]
void orc_get_margins(void* hv, int i, double* ls_margin, double* stop_margin) {
  Solver& s = S(hv, i);
  std::memcpy(ls_margin, s.ls_margin.data(), sizeof(double) * s.ls_margin.size());
  std::memcpy(stop_margin, s.stop_margin.data(), sizeof(double) * s.stop_margin.size());
}

// CPU baseline legs: every instance does one MPC step (initialize with u_init or warm start + solve) from its
// own x0. `threads` std::threads pull instances from a shared counter. Returns wall seconds.
double orc_mpc_step_batch(void* hv, const double* x0, const double* u_init, double* u_apply, double* cost_out, int threads) {
  Handle* h = static_cast<Handle*>(hv);
  int B = (int)h->solvers.size();
  if (threads < 1) threads = 1;
  auto t0 = std::chrono::steady_clock::now();
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < B; i = next.fetch_add(1))
      mpc_step(h->solvers[i], x0 + i * H1_NX, u_init, u_apply + i * H1_NU, cost_out ? cost_out + i : nullptr);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
int orc_max_threads(void) { unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

}  // extern "C"
