// Kernels built on the warp-cooperative f_D: batched one-step dynamics, nominal rollout with on-device cost,
// forward-difference linearization, and kinematic queries.
#pragma once
#include "h1_cost_eval.cuh"
#include "h1_lin_dirs.cuh"

namespace h1 {

// Stage the (read-only) model into shared memory; returns the aligned start of the per-warp scratch area.
template <class Model>
__device__ __forceinline__ unsigned char* stage_model(unsigned char* smem, const Model* g, const Model** out) {
  Model* s = reinterpret_cast<Model*>(smem);
  const int nwords = sizeof(Model) / 4;
  const unsigned* src = reinterpret_cast<const unsigned*>(g);
  unsigned* dst = reinterpret_cast<unsigned*>(s);
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  *out = s;
  return smem + ((sizeof(Model) + 15) / 16) * 16;
}

// ---- x_next[i] = f_D(x[i], u[i]) : one warp per state (RobotUtils::rolloutOneStep / step) ----
__global__ void k_dyn_step(const DynModel* gmd, int n, const double* __restrict__ x, const double* __restrict__ u,
                           double* __restrict__ xn) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  dyn_step_warp(*md, w, x + (size_t)i * NX, u + (size_t)i * NU, xn + (size_t)i * NX);
}

// ---- bias forces, dynamics-model CoM, ankle positions and sole contact points of arbitrary states
//      (computeGravComp, loadReferences' per-row FK, contact-schedule generation) ----
__global__ void k_dyn_query(const DynModel* gmd, int n, const double* __restrict__ x, double* __restrict__ bias,
                            double* __restrict__ com, double* __restrict__ ee, double* __restrict__ sole,
                            double* __restrict__ comvel, double* __restrict__ eevel) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  dyn_assemble_warp(*md, w, x + (size_t)i * NX, nullptr);
  if (bias && lane < NV) bias[(size_t)i * NV + lane] = w.biasv[lane];
  if (com && lane < 3) com[(size_t)i * 3 + lane] = w.com[lane];
  if (ee && lane < 6) ee[(size_t)i * 6 + lane] = w.q[lane % 3] + w.footr[lane / 3][lane % 3];
  if (sole && lane < NCPT)
    for (int c = 0; c < 3; ++c) sole[((size_t)i * NCPT + lane) * 3 + c] = w.q[c] + w.cp[lane][c];
  if (comvel) {
    // whole-body CoM velocity (world frame) = total linear momentum / total mass: the value of
    // mj_jacSubtreeCom(root) * qvel in RobotUtils::loadReferences (robot_utils.cpp:388-397). Lane b: spatial velocity
    // of body b about the base origin = sum of S_k v_k over its ancestor dofs, momentum l = m v_O - h x omega.
    double p[3] = {0.0, 0.0, 0.0}, m = 0.0;
    if (lane < NB) {
      const int j = 5 + lane, ns = md->nlist[j];
      double V[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int s = 0; s < ns; ++s) {
        const int k = md->alist[j][s];
        for (int c = 0; c < 6; ++c) V[c] += w.S[k][c] * w.v[k];
      }
      const double* I = &w.body[lane][6];
      m = I[0];
      double hw[3];
      cross3(I + 1, V, hw);
      for (int c = 0; c < 3; ++c) p[c] = m * V[3 + c] - hw[c];
    }
    m = warp_sum(m);
    for (int c = 0; c < 3; ++c) p[c] = warp_sum(p[c]);
    if (lane < 3) comvel[(size_t)i * 3 + lane] = p[lane] / m;
  }
  if (eevel && lane < H1_NFOOT) {
    // world velocity of the ankle body origin = jac_pos * qvel of RobotUtils::loadReferences (robot_utils.cpp:405-412):
    // spatial velocity of the foot body about the base origin, v = v_O + omega x r_foot
    const int j = md->foot_dof[lane], ns = md->nlist[j];
    double V[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int s = 0; s < ns; ++s) {
      const int k = md->alist[j][s];
      for (int c = 0; c < 6; ++c) V[c] += w.S[k][c] * w.v[k];
    }
    double wr[3];
    cross3(V, w.footr[lane], wr);
    for (int c = 0; c < 3; ++c) eevel[((size_t)i * H1_NFOOT + lane) * 3 + c] = V[3 + c] + wr[c];
  }
}

// ---- soft limit penalties of arbitrary (x, u) pairs with their derivatives: RobotUtils::constraintCost /
//      constraintGradients / constraintHessians (robot_utils.cpp:615-778). One thread per pair; u == nullptr: no control
//      terms (the terminal-knot form of terminalCost, robot_utils.cpp:226-250). Outputs may be nullptr. ----
__global__ void k_limit_penalties(const DynModel* gmd, const H1Weights* gw, int n, const double* __restrict__ x,
                                  const double* __restrict__ u, double* __restrict__ cost, double* __restrict__ gx,
                                  double* __restrict__ gu, double* __restrict__ hxx, double* __restrict__ huu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DynModel& md = *gmd;
  double c = 0.0;
  if (gx) for (int k = 0; k < NX; ++k) gx[(size_t)i * NX + k] = 0.0;
  if (hxx) for (int k = 0; k < NX; ++k) hxx[(size_t)i * NX + k] = 0.0;
  auto pen = [&](double val, double lo, double hi, double wgt, double* g, double* hh) {
    const double margin = 0.1 * (hi - lo), lo_s = lo + margin, hi_s = hi - margin;
    double gg = 0.0, h2 = 0.0;
    if (val > hi_s) { const double viol = val - hi_s; c += wgt * viol * viol; gg += 2.0 * wgt * viol; }
    if (val < lo_s) { const double viol = lo_s - val; c += wgt * viol * viol; gg += -2.0 * wgt * viol; }
    if (val > hi_s || val < lo_s) h2 = 2.0 * wgt;
    if (g) *g = gg;
    if (hh) *hh = h2;
  };
  for (int k = 0; k < NU; ++k) {
    if (u) pen(u[(size_t)i * NU + k], md.ctrl_lo[k], md.ctrl_hi[k], gw->w_control_limits, gu ? gu + (size_t)i * NU + k : nullptr,
               huu ? huu + (size_t)i * NU + k : nullptr);
    const double lo = md.jnt_lo[k], hi = md.jnt_hi[k];
    if (isfinite(lo) && isfinite(hi) && lo < hi)
      pen(x[(size_t)i * NX + 7 + k], lo, hi, gw->w_joint_limits, gx ? gx + (size_t)i * NX + 7 + k : nullptr,
          hxx ? hxx + (size_t)i * NX + 7 + k : nullptr);
  }
  if (cost) cost[i] = c;
}

// ---- RobotUtils::stageCost / terminalCost (robot_utils.cpp:162-252): 0.5 e'Qe + 0.5 eu'R eu + 0.5 w_com |com - com_ref|^2
//      + limit penalties for arbitrary states against given reference rows (diagonal Q / R / Qf). One warp per state;
//      u == nullptr: terminal form (Qf, joint limits only). ----
__global__ void k_stage_cost(const DynModel* gmd, const H1Weights* gw, int n, const double* __restrict__ x,
                             const double* __restrict__ u, const double* __restrict__ x_ref, const double* __restrict__ u_ref,
                             const double* __restrict__ com_ref, double* __restrict__ cost) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const bool terminal = u == nullptr;
  dyn_assemble_warp(*md, w, x + (size_t)i * NX, nullptr);
  const double* Qd = terminal ? gw->Qfdiag : gw->Qdiag;
  double acc = 0.0;
  for (int k = lane; k < NX; k += 32) { const double e = x[(size_t)i * NX + k] - x_ref[(size_t)i * NX + k]; acc += 0.5 * e * Qd[k] * e; }
  if (lane < NU) {
    if (!terminal) {
      const double uv = u[(size_t)i * NU + lane];
      const double e = uv - (u_ref ? u_ref[(size_t)i * NU + lane] : 0.0);
      acc += 0.5 * e * gw->Rdiag[lane] * e;
      acc += limit_pen(uv, md->ctrl_lo[lane], md->ctrl_hi[lane], gw->w_control_limits);
    }
    const double lo = md->jnt_lo[lane], hi = md->jnt_hi[lane];
    if (isfinite(lo) && isfinite(hi) && lo < hi) acc += limit_pen(x[(size_t)i * NX + 7 + lane], lo, hi, gw->w_joint_limits);
  }
  if (const double* qo = weights_offdiag(*gw)) {   // off-diagonal parts of full Q / R / Qf
    const double* Qo = qo + (terminal ? QOFF_QF : 0);
    for (int k = lane; k < NX; k += 32) {
      double sacc = 0.0;
      for (int j = 0; j < NX; ++j) sacc += Qo[j * NX + k] * (x[(size_t)i * NX + j] - x_ref[(size_t)i * NX + j]);
      acc += 0.5 * (x[(size_t)i * NX + k] - x_ref[(size_t)i * NX + k]) * sacc;
    }
    if (!terminal && lane < NU) {
      double sacc = 0.0;
      for (int j = 0; j < NU; ++j) sacc += qo[QOFF_R + j * NU + lane] * (u[(size_t)i * NU + j] - (u_ref ? u_ref[(size_t)i * NU + j] : 0.0));
      acc += 0.5 * (u[(size_t)i * NU + lane] - (u_ref ? u_ref[(size_t)i * NU + lane] : 0.0)) * sacc;
    }
  }
  if (lane < 3 && gw->w_com > 0.0 && com_ref) {
    const double e = w.com[lane] - com_ref[(size_t)i * 3 + lane];
    acc += 0.5 * gw->w_com * e * e;
  }
  acc = warp_sum(acc);
  if (lane == 0) cost[i] = acc;
}

// ---- nominal rollout xbar[t+1] = f_D(xbar[t], ubar[t]) with the trajectory cost as a by-product
//      (iLQR::forwardRolloutNominal + the baseline computeTotalCost of the following line search).
//      One warp per instance. t_begin > 0 rolls out only the tail (warm start: last knot). ----
__global__ void k_rollout(const DynModel* gmd, const H1Weights* gw, RefTable refs, int B, int N, int t_begin,
                          const int* __restrict__ active, const double* __restrict__ x0, double* __restrict__ xbar,
                          const double* __restrict__ ubar, double* __restrict__ cost_out,
                          PrimalFactor* __restrict__ pf_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (inst >= B) return;
  if (active && !active[inst]) return;
  const int lane = threadIdx.x & 31;
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  const double* ub = ubar + (size_t)inst * N * NU;
  if (x0) {
    for (int i = lane; i < NX; i += 32) xb[i] = x0[(size_t)inst * NX + i];
    __syncwarp();
  }
  const RefView r = refs.view(inst);
  double total = 0.0;
  for (int t = 0; t < N; ++t) {
    if (t >= t_begin) {
      if (pf_out) dyn_primal_factor_warp(*md, w, xb + t * NX, ub + t * NU, xb + (t + 1) * NX, pf_out[(size_t)inst * N + t]);
      else dyn_step_warp(*md, w, xb + t * NX, ub + t * NU, xb + (t + 1) * NX);
    } else if (cost_out) {
      dyn_assemble_warp(*md, w, xb + t * NX, ub + t * NU);
    }
    if (cost_out) total += knot_cost_warp(*md, w, *gw, r, t, ub + t * NU, false);
    __syncwarp();
  }
  if (cost_out) {
    dyn_assemble_warp(*md, w, xb + N * NX, nullptr);
    total += knot_cost_warp(*md, w, *gw, r, N, nullptr, true);
    if (lane == 0) cost_out[inst] = total;
  }
}

// ---- forward-difference linearization (iLQR::computeLinearization -> RobotUtils::linearizeDynamicsFD):
//      A[:,i] = (f(x + eps e_i, u) - f(x,u)) / eps, B[:,j] likewise; 1 + 51 + 19 evaluations per knot.
//      One CTA per (instance, knot); its warps share the 71 evaluations through a shared-memory table. ----
constexpr int LIN_WARPS = 6;
constexpr int LIN_EVALS = 1 + NX + NU;
__global__ void __launch_bounds__(LIN_WARPS * 32)
k_linearize_fd(const DynModel* gmd, int N, double eps, const int* __restrict__ active,
               const double* __restrict__ xbar, const double* __restrict__ ubar, double* __restrict__ A,
               double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int inst = blockIdx.x / N, t = blockIdx.x % N;
  if (active && !active[inst]) return;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp* ws = reinterpret_cast<DynWarp*>(p);
  double* fx = reinterpret_cast<double*>(p + sizeof(DynWarp) * LIN_WARPS);  // [LIN_EVALS][NX]
  double* xs = fx + LIN_EVALS * NX;                                         // [NX + NU] staged inputs
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const double* u = ubar + ((size_t)inst * N + t) * NU;
  for (int i = threadIdx.x; i < NX + NU; i += blockDim.x) xs[i] = (i < NX) ? x[i] : u[i - NX];
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  for (int e = warp; e < LIN_EVALS; e += LIN_WARPS)
    dyn_step_warp(*md, ws[warp], xs, xs + NX, fx + e * NX, e - 1, eps);
  __syncthreads();
  double* Ak = A + ((size_t)inst * N + t) * A_STRIDE;
  double* Bk = Bm + ((size_t)inst * N + t) * B_STRIDE;
  const double inv = 1.0 / eps;
  (void)inv;
  for (int i = threadIdx.x; i < NX * (NX + NU); i += blockDim.x) {
    const int col = i / NX, row = i - col * NX;
    const double v = (fx[(1 + col) * NX + row] - fx[row]) / eps;
    if (col < NX) Ak[i] = v; else Bk[i - NX * NX] = v;
  }
}

// ---- factorisation of Mhat at every knot of the current trajectory (granular API path; inside a solve the
//      nominal rollout produces the same factors as a by-product) ----
__global__ void k_primal_factor(const DynModel* gmd, int B, int N, const int* __restrict__ active,
                                const double* __restrict__ xbar, const double* __restrict__ ubar,
                                PrimalFactor* __restrict__ pf_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const long k = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= (long)B * N) return;
  const int inst = (int)(k / N), t = (int)(k % N);
  if (active && !active[inst]) return;
  dyn_primal_factor_warp(*md, w, xbar + ((size_t)inst * (N + 1) + t) * NX, ubar + ((size_t)inst * N + t) * NU, nullptr,
                         pf_out[k]);
}

// ---- analytic linearization: A = d f_D/dx, B = d f_D/du exactly. Per knot, ONE factorisation of Mhat (taken
//      from the nominal rollout) serves all 70 tangent directions; each direction is a dual-number
//      inverse-dynamics pass + two sparse triangular solves (h1_dyn.cuh). One CTA per (instance, knot),
//      directions round-robin over its warps, columns written straight to A_k / B_k. ----
constexpr int LINA_WARPS = 4;
__global__ void __launch_bounds__(LINA_WARPS * 32, 3)
k_linearize_analytic(const DynModel* gmd, int N, const int* __restrict__ active, const double* __restrict__ xbar,
                     const double* __restrict__ ubar, const PrimalFactor* __restrict__ pf_g, double* __restrict__ A,
                     double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int inst = blockIdx.x / N, t = blockIdx.x % N;
  if (active && !active[inst]) return;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  TanWarpT<Dual>* ws = reinterpret_cast<TanWarpT<Dual>*>(p);
  PrimalFactor* pf = reinterpret_cast<PrimalFactor*>(p + sizeof(TanWarpT<Dual>) * LINA_WARPS);
  double* xs = reinterpret_cast<double*>(pf + 1);
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const double* u = ubar + ((size_t)inst * N + t) * NU;
  for (int i = threadIdx.x; i < NX + NU; i += blockDim.x) xs[i] = (i < NX) ? x[i] : u[i - NX];
  {
    const double* src = reinterpret_cast<const double*>(pf_g + ((size_t)inst * N + t));
    double* dst = reinterpret_cast<double*>(pf);
    for (int i = threadIdx.x; i < (int)(sizeof(PrimalFactor) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  double* Ak = A + ((size_t)inst * N + t) * A_STRIDE;
  double* Bk = Bm + ((size_t)inst * N + t) * B_STRIDE;
  for (int e = warp; e < NX + NU; e += LINA_WARPS)
    dyn_tangent_id_warp(*md, ws[warp], *pf, xs, xs + NX, e, e < NX ? Ak + e * NX : Bk + (e - NX) * NX);
}

// ---- analytic linearization, one THREAD per column of [A_k | B_k] (h1_lin_dirs.cuh). Threads are packed
//      (knot, direction) -> global thread id, so warps are full regardless of the direction count; the three
//      direction classes are separate instantiations (launches) with their own register budgets.
//      MODE 0: the 26 q columns, 1: the 25 v columns, 2: the 19 control columns. ----
constexpr int LIND_THREADS = 128;
template <int MODE, bool H1TREE>   // H1TREE: the model has H1's dof tree (DynModel::seq_ok) -> static-index solve
__global__ void __launch_bounds__(LIND_THREADS)
k_linearize_dirs(const DynModel* gmd, long nknots, int N, const int* __restrict__ active,
                 const double* __restrict__ xbar, const double* __restrict__ ubar,
                 const PrimalFactor* __restrict__ pf_g, double* __restrict__ A, double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  stage_model(smem, gmd, &md);
  constexpr int ND = MODE == 0 ? NQ : (MODE == 1 ? NV : NU);
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long knot = g / ND;
  const int dir = (int)(g - knot * ND);
  if (knot >= nknots) return;
  const long inst = knot / N;
  const int t = (int)(knot - inst * N);
  if (active && !active[inst]) return;
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const PrimalFactor* pf = pf_g + knot;
  double tv[NV];
  int seed;
  if (MODE == 0) { seed = dir; id_tangent_seq<Dual, Dual>(*md, x, pf->a, seed, tv); }
  else if (MODE == 1) { seed = NQ + dir; id_tangent_seq<double, Dual>(*md, x, pf->a, seed, tv); }
  else {
    seed = NX + dir;
    const double uj = ubar[(size_t)knot * NU + dir];
    for (int j = 0; j < NV; ++j) tv[j] = 0.0;
    tv[6 + dir] = (uj < md->ctrl_lo[dir] || uj > md->ctrl_hi[dir]) ? 0.0 : 1.0;   // clamped torque: no sensitivity
  }
  if (H1TREE) tangent_solve_h1(&pf->Lm[0][0], pf->D, tv);
  else tangent_solve_seq(*md, &pf->Lm[0][0], pf->D, tv);
  double* col = (MODE == 2) ? Bm + (size_t)knot * B_STRIDE + (size_t)dir * NX : A + (size_t)knot * A_STRIDE + (size_t)seed * NX;
  integrate_tangent_seq(*md, x, pf->a, seed, tv, col);
}

}  // namespace h1
