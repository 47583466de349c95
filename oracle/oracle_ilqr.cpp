// TEST INFRASTRUCTURE ONLY — CPU oracle of the iLQR control flow, Riccati recursion, line search and MPC
// step, restated behaviour-for-behaviour (including the quirks listed in SURVEY.md Appendix B) from
//   iLQR::initializeWithReference  /root/reference/src/ilqr/ilqr.cpp:50-117
//   iLQR::forwardRolloutNominal    ilqr.cpp:119-124
//   iLQR::computeLinearization     ilqr.cpp:126-131
//   iLQR::computeCostQuadratics    ilqr.cpp:133-244   (+ add*CostDerivatives ilqr.cpp:662-800)
//   iLQR::backwardPass             ilqr.cpp:250-309
//   iLQR::forwardPassLineSearch    ilqr.cpp:311-361
//   iLQR::computeTotalCost         ilqr.cpp:363-518
//   iLQR::solve                    ilqr.cpp:521-660
//   MPC::stepOnce                  /root/reference/src/ilqr/mpc.cpp:40-127
// Parity status: UNPINNED (no reference tests / golden vectors exist). Eigen's LLT / LDLT are restated as a
// plain Cholesky positive-definiteness test and an LDL^T with largest-|diagonal| symmetric pivoting.
#include "oracle.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

static const int NX = H1_NX, NU = H1_NU;

void Solver::init(Problem* prob) {
  p = prob; N = prob->N; lambda = prob->opt.reg_init;
  xbar.assign((N + 1) * NX, 0.0); ubar.assign(N * NU, 0.0);
  K.assign(N * NU * NX, 0.0); kff.assign(N * NU, 0.0);
  A.assign(N * NX * NX, 0.0); B.assign(N * NX * NU, 0.0);
  lx.assign((N + 1) * NX, 0.0); lu.assign(N * NU, 0.0);
  lxx.assign((N + 1) * NX * NX, 0.0); luu.assign(N * NU * NU, 0.0);
  has_prev = false; iters = 0;
}

void rollout_nominal(Solver& s, const double* x0) {
  std::memcpy(&s.xbar[0], x0, sizeof(double) * NX);
  for (int t = 0; t < s.N; ++t) dyn_step(s.p->dyn, &s.xbar[t * NX], &s.ubar[t * NU], &s.xbar[(t + 1) * NX]);
}

void linearize(Solver& s) {
  for (int t = 0; t < s.N; ++t) {
    if (s.p->opt.linearization == H1ILQR_LIN_FD)
      dyn_linearize_fd(s.p->dyn, &s.xbar[t * NX], &s.ubar[t * NU], s.p->opt.fd_eps, &s.A[t * NX * NX], &s.B[t * NX * NU]);
    else
      dyn_linearize_ad(s.p->dyn, &s.xbar[t * NX], &s.ubar[t * NU], &s.A[t * NX * NX], &s.B[t * NX * NU]);
  }
}

static void term(const Problem& p, int t, int ee, const double* x, const double* target, double w, double* g, double* H) {
  if (p.use_ad) cost_term_ad(p.cost, t, ee, x, target, w, g, H);
  else cost_term_analytic(p.cost, t, ee, x, target, w, g, H);
}

// support centre from horizon-local stance flags and foot targets (ilqr.cpp:763-791, quirk Q6)
static bool support_centre(const Problem& p, int t, double* ps) {
  bool ls = p.stance[2 * t] == 1, rs = p.stance[2 * t + 1] == 1;
  const double* l = &p.ee_ref[6 * t]; const double* r = &p.ee_ref[6 * t + 3];
  if (ls && rs) { ps[0] = 0.5 * (l[0] + r[0]); ps[1] = 0.5 * (l[1] + r[1]); }
  else if (ls) { ps[0] = l[0]; ps[1] = l[1]; }
  else if (rs) { ps[0] = r[0]; ps[1] = r[1]; }
  else return false;
  return true;
}

static void kinematic_terms(const Problem& p, int t, bool terminal, const double* x, double* g, double* H) {
  const H1Weights& w = p.w;
  if (w.w_com > 0.0) term(p, TERM_COM, 0, x, &p.com_ref[3 * t], w.w_com, g, H);
  if (!terminal && w.w_com_vel > 0.0) term(p, TERM_COM_VEL, 0, x, &p.com_vel_ref[3 * t], w.w_com_vel, g, H);
  if (w.w_ee_pos > 0.0)
    for (int ee = 0; ee < 2; ++ee)
      if (p.stance[2 * t + ee] != 1) term(p, TERM_EE_POS, ee, x, &p.ee_ref[6 * t + 3 * ee], w.w_ee_pos, g, H);
  if (w.w_ee_vel > 0.0) {
    const double zero[3] = {0, 0, 0};
    for (int ee = 0; ee < 2; ++ee)
      if (p.stance[2 * t + ee] == 1) term(p, TERM_EE_VEL, ee, x, zero, w.w_ee_vel, g, H);
  }
  if (w.w_upright > 0.0) term(p, TERM_UPRIGHT, 0, x, nullptr, w.w_upright, g, H);
  if (w.w_balance > 0.0) {
    double ps[2];
    if (support_centre(p, t, ps)) term(p, TERM_BALANCE, 0, x, ps, w.w_balance, g, H);
  }
}

void cost_quadratics(Solver& s) {
  const Problem& p = *s.p;
  const int N = s.N;
  double gx[NX], gu[NU], hx[NX], hu[NU];
  for (int t = 0; t < N; ++t) {
    const double* x = &s.xbar[t * NX]; const double* u = &s.ubar[t * NU];
    double* lx = &s.lx[t * NX]; double* lu = &s.lu[t * NU];
    double* lxx = &s.lxx[t * NX * NX]; double* luu = &s.luu[t * NU * NU];
    std::fill(lxx, lxx + NX * NX, 0.0); std::fill(luu, luu + NU * NU, 0.0);
    if (p.Qfull.empty()) {
      for (int i = 0; i < NX; ++i) { lx[i] = p.w.Qdiag[i] * (x[i] - p.x_ref[t * NX + i]); lxx[i * NX + i] = p.w.Qdiag[i]; }
      for (int i = 0; i < NU; ++i) { lu[i] = p.w.Rdiag[i] * (u[i] - p.u_ref[t * NU + i]); luu[i * NU + i] = p.w.Rdiag[i]; }
    } else {   // lx = Q (x - x_ref), lxx = Q, lu = R (u - u_ref), luu = R with whole matrices (ilqr.cpp:145-150)
      for (int i = 0; i < NX; ++i) {
        double a = 0;
        for (int j = 0; j < NX; ++j) { a += p.Qfull[j * NX + i] * (x[j] - p.x_ref[t * NX + j]); lxx[j * NX + i] = p.Qfull[j * NX + i]; }
        lx[i] = a;
      }
      for (int i = 0; i < NU; ++i) {
        double a = 0;
        for (int j = 0; j < NU; ++j) { a += p.Rfull[j * NU + i] * (u[j] - p.u_ref[t * NU + j]); luu[j * NU + i] = p.Rfull[j * NU + i]; }
        lu[i] = a;
      }
    }
    kinematic_terms(p, t, false, x, lx, lxx);
    limit_derivs(p.dyn, p.w, x, u, gx, gu, hx, hu);
    for (int i = 0; i < NX; ++i) { lx[i] += gx[i]; lxx[i * NX + i] += hx[i]; }
    for (int i = 0; i < NU; ++i) { lu[i] += gu[i]; luu[i * NU + i] += hu[i]; }
  }
  const double* x = &s.xbar[N * NX];
  double* lx = &s.lx[N * NX]; double* lxx = &s.lxx[N * NX * NX];
  std::fill(lxx, lxx + NX * NX, 0.0);
  if (p.Qffull.empty()) {
    for (int i = 0; i < NX; ++i) { lx[i] = p.w.Qfdiag[i] * (x[i] - p.x_ref[N * NX + i]); lxx[i * NX + i] = p.w.Qfdiag[i]; }
  } else {
    for (int i = 0; i < NX; ++i) {
      double a = 0;
      for (int j = 0; j < NX; ++j) { a += p.Qffull[j * NX + i] * (x[j] - p.x_ref[N * NX + j]); lxx[j * NX + i] = p.Qffull[j * NX + i]; }
      lx[i] = a;
    }
  }
  kinematic_terms(p, N, true, x, lx, lxx);
  double uz[NU] = {0};
  limit_derivs(p.dyn, p.w, x, uz, gx, gu, hx, hu);
  for (int i = 0; i < NX; ++i) { lx[i] += gx[i]; lxx[i * NX + i] += hx[i]; }
}

// ---- dense helpers, column-major ----
static void gemm_tn(int m, int n, int k, const double* A, int lda, const double* Bm, int ldb, double* C, int ldc) {
  // C(m x n) = A^T (A is k x m) * B (k x n)
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double acc = 0;
      for (int l = 0; l < k; ++l) acc += A[i * lda + l] * Bm[j * ldb + l];
      C[j * ldc + i] = acc;
    }
}
static void gemm_nn(int m, int n, int k, const double* A, int lda, const double* Bm, int ldb, double* C, int ldc) {
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) C[j * ldc + i] = 0;
    for (int l = 0; l < k; ++l) {
      double b = Bm[j * ldb + l];
      for (int i = 0; i < m; ++i) C[j * ldc + i] += A[l * lda + i] * b;
    }
  }
}

static bool cholesky_ok(const double* Q, int n) {  // Eigen::LLT info() == Success
  double L[NU * NU];
  std::memcpy(L, Q, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double d = L[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[k * n + j] * L[k * n + j];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double sacc = L[j * n + i];
      for (int k = 0; k < j; ++k) sacc -= L[k * n + i] * L[k * n + j];
      L[j * n + i] = sacc / d;
    }
  }
  return true;
}

// LDL^T with symmetric pivoting on the largest |diagonal| (Eigen::LDLT), then solve Q X = RHS (n x nrhs)
static void ldlt_solve(const double* Q, int n, double* X, int nrhs) {
  double M[NU * NU];
  int perm[NU];
  std::memcpy(M, Q, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  // order: at step k pick the remaining index with the largest |original diagonal| (left-looking LDLT
  // only updates the diagonal of the pivot row before using it)
  for (int k = 0; k < n; ++k) {
    int best = k;
    double bv = std::fabs(Q[perm[k] * n + perm[k]]);
    for (int i = k + 1; i < n; ++i) { double v = std::fabs(Q[perm[i] * n + perm[i]]); if (v > bv) { bv = v; best = i; } }
    std::swap(perm[k], perm[best]);
  }
  double Pm[NU * NU], D[NU];
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) Pm[j * n + i] = Q[perm[j] * n + perm[i]];
  // L (unit lower) stored in Pm's lower part
  for (int k = 0; k < n; ++k) {
    double d = Pm[k * n + k];
    for (int c = 0; c < k; ++c) d -= Pm[c * n + k] * Pm[c * n + k] * D[c];
    D[k] = d;
    for (int i = k + 1; i < n; ++i) {
      double sacc = Pm[k * n + i];
      for (int c = 0; c < k; ++c) sacc -= Pm[c * n + i] * Pm[c * n + k] * D[c];
      Pm[k * n + i] = (std::fabs(d) > 0.0) ? sacc / d : 0.0;
    }
  }
  const double tol = 2.2250738585072014e-308;  // numeric_limits<double>::min(), Eigen LDLT::_solve_impl
  double y[NU];
  for (int r = 0; r < nrhs; ++r) {
    double* x = X + r * n;
    for (int i = 0; i < n; ++i) y[i] = x[perm[i]];
    for (int i = 0; i < n; ++i) for (int c = 0; c < i; ++c) y[i] -= Pm[c * n + i] * y[c];
    for (int i = 0; i < n; ++i) y[i] = (std::fabs(D[i]) > tol) ? y[i] / D[i] : 0.0;
    for (int i = n - 1; i >= 0; --i) for (int c = i + 1; c < n; ++c) y[i] -= Pm[i * n + c] * y[c];
    for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
  }
  (void)M;
}

void backward_pass(Solver& s) {
  const int N = s.N;
  std::vector<double> Vx(&s.lx[N * NX], &s.lx[N * NX] + NX), Vxx(&s.lxx[N * NX * NX], &s.lxx[N * NX * NX] + NX * NX);
  std::vector<double> VA(NX * NX), VB(NX * NU), Qxx(NX * NX), Quu(NU * NU), Qxu(NX * NU), Qx(NX), Qu(NU);
  std::vector<double> QuuK(NU * NX), T1(NX * NX), T2(NX * NX), T3(NX * NX), Kt_rhs(NU * NX);
  for (int t = N - 1; t >= 0; --t) {
    const double* A = &s.A[t * NX * NX]; const double* B = &s.B[t * NX * NU];
    double* K = &s.K[t * NU * NX]; double* k = &s.kff[t * NU];
    for (int i = 0; i < NX; ++i) { double a = 0; for (int l = 0; l < NX; ++l) a += A[i * NX + l] * Vx[l]; Qx[i] = s.lx[t * NX + i] + a; }
    for (int i = 0; i < NU; ++i) { double a = 0; for (int l = 0; l < NX; ++l) a += B[i * NX + l] * Vx[l]; Qu[i] = s.lu[t * NU + i] + a; }
    gemm_nn(NX, NX, NX, Vxx.data(), NX, A, NX, VA.data(), NX);
    gemm_nn(NX, NU, NX, Vxx.data(), NX, B, NX, VB.data(), NX);
    gemm_tn(NX, NX, NX, A, NX, VA.data(), NX, Qxx.data(), NX);
    gemm_tn(NU, NU, NX, B, NX, VB.data(), NX, Quu.data(), NU);
    gemm_tn(NX, NU, NX, A, NX, VB.data(), NX, Qxu.data(), NX);  // lxu == 0 (ilqr.cpp:151)
    for (int i = 0; i < NX * NX; ++i) Qxx[i] += s.lxx[t * NX * NX + i];
    for (int i = 0; i < NU * NU; ++i) Quu[i] += s.luu[t * NU * NU + i];
    for (int i = 0; i < NU; ++i) Quu[i * NU + i] += s.lambda;
    if (!cholesky_ok(Quu.data(), NU))
      for (int i = 0; i < NU; ++i) Quu[i * NU + i] += 1e-4;  // once, no re-check (Q9)
    // K = -Quu^{-1} Qxu^T (19x51), k = -Quu^{-1} Qu
    for (int j = 0; j < NX; ++j) for (int i = 0; i < NU; ++i) Kt_rhs[j * NU + i] = Qxu[i * NX + j];
    ldlt_solve(Quu.data(), NU, Kt_rhs.data(), NX);
    for (int i = 0; i < NU * NX; ++i) K[i] = -Kt_rhs[i];
    for (int i = 0; i < NU; ++i) k[i] = Qu[i];
    ldlt_solve(Quu.data(), NU, k, 1);
    for (int i = 0; i < NU; ++i) k[i] = -k[i];
    // Vx = Qx + K^T Quu k + K^T Qu + Qxu k
    double Quuk[NU];
    for (int i = 0; i < NU; ++i) { double a = 0; for (int l = 0; l < NU; ++l) a += Quu[l * NU + i] * k[l]; Quuk[i] = a; }
    for (int i = 0; i < NX; ++i) {
      double a1 = 0, a2 = 0, a3 = 0;
      for (int l = 0; l < NU; ++l) { a1 += K[i * NU + l] * Quuk[l]; a2 += K[i * NU + l] * Qu[l]; a3 += Qxu[l * NX + i] * k[l]; }
      Vx[i] = Qx[i] + a1 + a2 + a3;
    }
    // Vxx = Qxx + K^T Quu K + K^T Qxu^T + Qxu K ; then symmetrise
    gemm_nn(NU, NX, NU, Quu.data(), NU, K, NU, QuuK.data(), NU);
    gemm_tn(NX, NX, NU, K, NU, QuuK.data(), NU, T1.data(), NX);
    gemm_nn(NX, NX, NU, Qxu.data(), NX, K, NU, T3.data(), NX);  // Qxu K
    for (int j = 0; j < NX; ++j) for (int i = 0; i < NX; ++i) T2[j * NX + i] = T3[i * NX + j];  // K^T Qxu^T = (Qxu K)^T
    for (int i = 0; i < NX * NX; ++i) T1[i] = Qxx[i] + T1[i] + T2[i] + T3[i];
    for (int j = 0; j < NX; ++j) for (int i = 0; i < NX; ++i) Vxx[j * NX + i] = 0.5 * (T1[j * NX + i] + T1[i * NX + j]);
  }
}

double total_cost(const Solver& s, const double* xt, const double* ut) {
  const Problem& p = *s.p;
  const int N = s.N;
  double total = 0.0;
  auto upright_balance = [&](int t, const double* x) {
    if (p.w.w_upright > 0.0) {  // MuJoCo roles here (ilqr.cpp:380-397), unlike the derivative (Q4)
      double qw = x[3], qx = x[4], qy = x[5], qz = x[6];
      double z0 = 2.0 * (qx * qz + qw * qy), z1 = 2.0 * (qy * qz - qw * qx), z2 = 1.0 - 2.0 * (qx * qx + qy * qy) - 1.0;
      total += 0.5 * p.w.w_upright * (z0 * z0 + z1 * z1 + z2 * z2);
    }
    if (p.w.w_balance > 0.0) {
      double ps[2];
      if (support_centre(p, t, ps)) {
        double com[3];
        dyn_com(p.dyn, x, com);  // dynamics-model CoM and raw base linear velocity (Q7)
        double om = std::sqrt(com[2] / 9.81);
        double r0 = com[0] + x[H1_NQ] * om - ps[0], r1 = com[1] + x[H1_NQ + 1] * om - ps[1];
        total += 0.5 * p.w.w_balance * (r0 * r0 + r1 * r1);
      }
    }
  };
  for (int t = 0; t < N; ++t) {
    const double* x = xt + t * NX; const double* u = ut + t * NU;
    double qx = 0, qu = 0;
    if (p.Qfull.empty()) {
      for (int i = 0; i < NX; ++i) { double e = x[i] - p.x_ref[t * NX + i]; qx += e * p.w.Qdiag[i] * e; }
      for (int i = 0; i < NU; ++i) { double e = u[i] - p.u_ref[t * NU + i]; qu += e * p.w.Rdiag[i] * e; }
    } else {
      for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j)
        qx += (x[i] - p.x_ref[t * NX + i]) * p.Qfull[j * NX + i] * (x[j] - p.x_ref[t * NX + j]);
      for (int i = 0; i < NU; ++i) for (int j = 0; j < NU; ++j)
        qu += (u[i] - p.u_ref[t * NU + i]) * p.Rfull[j * NU + i] * (u[j] - p.u_ref[t * NU + j]);
    }
    total += 0.5 * qx; total += 0.5 * qu;
    upright_balance(t, x);
  }
  const double* xN = xt + N * NX;
  double qf = 0;
  if (p.Qffull.empty()) {
    for (int i = 0; i < NX; ++i) { double e = xN[i] - p.x_ref[N * NX + i]; qf += e * p.w.Qfdiag[i] * e; }
  } else {
    for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j)
      qf += (xN[i] - p.x_ref[N * NX + i]) * p.Qffull[j * NX + i] * (xN[j] - p.x_ref[N * NX + j]);
  }
  total += 0.5 * qf;
  upright_balance(N, xN);
  for (int t = 0; t < N; ++t) total += limit_cost(p.dyn, p.w, xt + t * NX, ut + t * NU);
  double uz[NU] = {0};
  total += limit_cost(p.dyn, p.w, xN, uz);
  return total;
}

bool line_search(Solver& s, const double* x0, double* new_cost, int* alpha_index) {
  const Problem& p = *s.p;
  const int N = s.N;
  double baseline = total_cost(s, s.xbar.data(), s.ubar.data());
  s.last_ls_margin = -1.0;
  std::vector<double> xn((N + 1) * NX), un(N * NU);
  for (int ai = 0; ai < H1ILQR_NALPHA; ++ai) {
    double alpha = p.opt.alphas[ai];
    std::memcpy(&xn[0], x0, sizeof(double) * NX);
    for (int t = 0; t < N; ++t) {
      const double* K = &s.K[t * NU * NX];
      for (int i = 0; i < NU; ++i) {
        double acc = 0;
        for (int l = 0; l < NX; ++l) acc += K[l * NU + i] * (xn[t * NX + l] - s.xbar[t * NX + l]);
        un[t * NU + i] = s.ubar[t * NU + i] + alpha * s.kff[t * NU + i] + acc;
      }
      dyn_step(p.dyn, &xn[t * NX], &un[t * NU], &xn[(t + 1) * NX]);
    }
    double c = total_cost(s, xn.data(), un.data());
    {
      // a non-finite candidate cost is rejected whatever the rounding: it does not narrow the margin
      const double m = std::fabs(c - (baseline - p.opt.accept_margin));
      if (m == m && (s.last_ls_margin < 0.0 || m < s.last_ls_margin)) s.last_ls_margin = m;
    }
    if (c < baseline - p.opt.accept_margin) {
      s.xbar = xn; s.ubar = un;
      *new_cost = c; *alpha_index = ai;
      return true;
    }
  }
  *new_cost = baseline; *alpha_index = -1;
  return false;
}

void initialize(Solver& s, const double* x0, bool warm, const double* u_init) {
  const int N = s.N;
  std::memcpy(&s.xbar[0], x0, sizeof(double) * NX);
  if (warm && s.has_prev) {
    for (int t = 0; t < N - 1; ++t) std::memcpy(&s.ubar[t * NU], &s.prev_ubar[(t + 1) * NU], sizeof(double) * NU);
    std::memcpy(&s.ubar[(N - 1) * NU], &s.prev_ubar[(N - 1) * NU], sizeof(double) * NU);
    for (int t = 0; t < N - 1; ++t) std::memcpy(&s.xbar[(t + 1) * NX], &s.prev_xbar[(t + 2) * NX], sizeof(double) * NX);
    dyn_step(s.p->dyn, &s.xbar[(N - 1) * NX], &s.ubar[(N - 1) * NU], &s.xbar[N * NX]);
  } else {
    for (int t = 0; t < N; ++t) std::memcpy(&s.ubar[t * NU], u_init, sizeof(double) * NU);
    for (int t = 0; t < N; ++t) dyn_step(s.p->dyn, &s.xbar[t * NX], &s.ubar[t * NU], &s.xbar[(t + 1) * NX]);
  }
}

bool solve(Solver& s, const double* x0, double* cost_out) {
  const H1SolverOptions& o = s.p->opt;
  s.cost_trace.assign(o.max_iterations, 0.0);
  s.alpha_trace.assign(2 * o.max_iterations, -2);
  s.ls_margin.assign(2 * o.max_iterations, -1.0);
  s.stop_margin.assign(o.max_iterations, -1.0);
  double cur = total_cost(s, s.xbar.data(), s.ubar.data());
  int it = 0;
  for (it = 0; it < o.max_iterations; ++it) {
    double prev = cur;
    rollout_nominal(s, x0);
    linearize(s);
    cost_quadratics(s);
    backward_pass(s);
    double nc; int ai;
    bool improved = line_search(s, x0, &nc, &ai);
    s.alpha_trace[2 * it] = ai;
    s.ls_margin[2 * it] = s.last_ls_margin;
    if (!improved) {
      s.lambda = std::min(s.lambda * 10.0, o.reg_max);
      backward_pass(s);
      improved = line_search(s, x0, &nc, &ai);
      s.alpha_trace[2 * it + 1] = ai;
      s.ls_margin[2 * it + 1] = s.last_ls_margin;
      if (!improved) {
        s.cost_trace[it] = cur;
        if (it > 1) { ++it; break; }
        continue;
      }
    }
    cur = nc;
    s.lambda = std::max(s.lambda / 2.0, o.reg_min);
    s.cost_trace[it] = cur;
    s.stop_margin[it] = std::min(std::fabs(std::fabs(cur - prev) - o.tolerance), std::fabs(cur - o.divergence_cost));
    if (std::fabs(cur - prev) < o.tolerance) { ++it; break; }
    if (cur > o.divergence_cost) { ++it; break; }
  }
  s.iters = it;
  *cost_out = cur;
  return true;
}

bool mpc_step(Solver& s, const double* x_meas, const double* u_init, double* u_apply, double* cost_out) {
  initialize(s, x_meas, s.has_prev, u_init);
  double c;
  solve(s, x_meas, &c);
  for (int i = 0; i < NU; ++i) {
    double acc = 0;
    for (int l = 0; l < NX; ++l) acc += s.K[l * NU + i] * (x_meas[l] - s.xbar[l]);
    u_apply[i] = s.ubar[i] + acc;
  }
  s.prev_xbar = s.xbar; s.prev_ubar = s.ubar; s.has_prev = true;
  if (cost_out) *cost_out = c;
  return true;
}

}  // namespace orc
