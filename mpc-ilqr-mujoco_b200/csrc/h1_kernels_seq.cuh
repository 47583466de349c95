// Throughput-form kernels built on the thread-sequential f_D (h1_dyn_seq.cuh): nominal rollout with one thread
// per instance and the alpha-parallel line search with one thread per (instance, candidate). They compute the
// same stages as k_rollout / k_line_search (h1_kernels_dyn.cuh, h1_kernels_solve.cuh) with a different
// operation order; the C ABI picks a family per H1ILQR_KERNELS_* policy.
#pragma once
#include "h1_cost_eval.cuh"
#include "h1_dyn_seq.cuh"
#include "h1_kernels_dyn.cuh"

namespace h1 {

// Cost of knot t for state x (raw 51), control u (nullptr at the terminal knot) and dynamics-model CoM `com`:
// iLQR::computeTotalCost / RobotUtils::constraintCost, same terms and quirks as knot_cost_warp.
__device__ __forceinline__ double knot_cost_seq(const DynModel& md, const H1Weights& wt, const RefView& r, int t,
                                                const double* __restrict__ x, const double* __restrict__ u,
                                                const double* com, bool terminal) {
  const double* Qd = terminal ? wt.Qfdiag : wt.Qdiag;
  const double* xr = r.x_ref + t * NX;
  double acc = 0.0;
#pragma unroll 1
  for (int i = 0; i < NX; ++i) {
    const double e = x[i] - xr[i];
    acc += 0.5 * e * Qd[i] * e;
  }
  if (const double* qo = weights_offdiag(wt)) {   // off-diagonal parts of full Q / R / Qf
    const double* Qo = qo + (terminal ? QOFF_QF : 0);
#pragma unroll 1
    for (int i = 0; i < NX; ++i) {
      double sacc = 0.0;
#pragma unroll 1
      for (int j = 0; j < NX; ++j) sacc += Qo[j * NX + i] * (x[j] - xr[j]);
      acc += 0.5 * (x[i] - xr[i]) * sacc;
    }
    if (!terminal) {
#pragma unroll 1
      for (int i = 0; i < NU; ++i) {
        double sacc = 0.0;
#pragma unroll 1
        for (int j = 0; j < NU; ++j) sacc += qo[QOFF_R + j * NU + i] * (u[j] - r.u_ref[t * NU + j]);
        acc += 0.5 * (u[i] - r.u_ref[t * NU + i]) * sacc;
      }
    }
  }
#pragma unroll 1
  for (int i = 0; i < NU; ++i) {
    const double ui = terminal ? 0.0 : u[i];
    if (!terminal) { const double e = ui - r.u_ref[t * NU + i]; acc += 0.5 * e * wt.Rdiag[i] * e; }
    acc += limit_pen(ui, md.ctrl_lo[i], md.ctrl_hi[i], wt.w_control_limits);
    const double lo = md.jnt_lo[i], hi = md.jnt_hi[i];
    if (isfinite(lo) && isfinite(hi) && lo < hi) acc += limit_pen(x[7 + i], lo, hi, wt.w_joint_limits);
  }
  if (wt.w_upright > 0.0) {
    const double qw = x[3], qx = x[4], qy = x[5], qz = x[6];
    const double z0 = 2.0 * (qx * qz + qw * qy), z1 = 2.0 * (qy * qz - qw * qx);
    const double z2 = (1.0 - 2.0 * (qx * qx + qy * qy)) - 1.0;
    acc += 0.5 * wt.w_upright * (z0 * z0 + z1 * z1 + z2 * z2);
  }
  if (wt.w_balance > 0.0) {
    double ps[2];
    if (support_centre(r, t, ps)) {
      const double om = sqrt(com[2] / 9.81);
      const double r0 = com[0] + x[NQ] * om - ps[0], r1 = com[1] + x[NQ + 1] * om - ps[1];
      acc += 0.5 * wt.w_balance * (r0 * r0 + r1 * r1);
    }
  }
  return acc;
}

// ---- nominal rollout, one thread per instance (iLQR::forwardRolloutNominal + baseline computeTotalCost);
//      same contract as k_rollout ----
#ifndef H1_SEQ_ROLL_THREADS
#define H1_SEQ_ROLL_THREADS 32
#endif
constexpr int SEQ_ROLL_THREADS = H1_SEQ_ROLL_THREADS;   // (128 measured the same: these kernels are latency bound per warp)
__global__ void __launch_bounds__(SEQ_ROLL_THREADS)
k_rollout_seq(const DynModel* gmd, const H1Weights* gw, RefTable refs, int B, int N, int t_begin,
              const int* __restrict__ active, const double* __restrict__ x0, double* __restrict__ xbar,
              const double* __restrict__ ubar, double* __restrict__ cost_out, PrimalFactor* __restrict__ pf_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  stage_model(smem, gmd, &md);
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= B) return;
  if (active && !active[inst]) return;
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  const double* ub = ubar + (size_t)inst * N * NU;
  if (x0)
    for (int i = 0; i < NX; ++i) xb[i] = x0[(size_t)inst * NX + i];
  const RefView r = refs.view(inst);
  double total = 0.0, com[3];
#pragma unroll 1
  for (int t = 0; t < N; ++t) {
    if (t >= t_begin)
      dyn_step_seq(*md, xb + t * NX, ub + t * NU, xb + (t + 1) * NX, pf_out ? pf_out + (size_t)inst * N + t : nullptr,
                   cost_out ? com : nullptr);
    else if (cost_out)
      dyn_com_seq(*md, xb + t * NX, com);
    if (cost_out) total += knot_cost_seq(*md, *gw, r, t, xb + t * NX, ub + t * NU, com, false);
  }
  if (cost_out) {
    dyn_com_seq(*md, xb + N * NX, com);
    total += knot_cost_seq(*md, *gw, r, N, xb + N * NX, nullptr, com, true);
    cost_out[inst] = total;
  }
}

// ---- factorisation of Mhat at every knot of the current trajectory, one thread per (instance, knot): the
//      knot-parallel replacement of the nominal rollout in iterations >= 1, where the rollout would only
//      reproduce the trajectory the line search has just accepted (f_D is deterministic) ----
__global__ void __launch_bounds__(SEQ_ROLL_THREADS)
k_primal_factor_seq(const DynModel* gmd, int B, int N, const int* __restrict__ active, const int* __restrict__ list,
                    const int* __restrict__ list_count, const double* __restrict__ xbar,
                    const double* __restrict__ ubar, PrimalFactor* __restrict__ pf_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
  // `list` (optional): compact list of the active instances (k_solve_state): full warps however sparse the set is
  if (list && (long)blockIdx.x * blockDim.x >= (long)(*list_count) * N) return;
  const DynModel* md;
  stage_model(smem, gmd, &md);
  if (s >= (long)B * N) return;
  int inst = (int)(s / N);
  const int t = (int)(s - (long)inst * N);
  if (list) { if (inst >= *list_count) return; inst = list[inst]; }
  else if (active && !active[inst]) return;
  dyn_step_seq(*md, xbar + ((size_t)inst * (N + 1) + t) * NX, ubar + ((size_t)inst * N + t) * NU, nullptr,
               pf_out + (size_t)inst * N + t, nullptr);
}

// ---- line search, one thread per (instance, alpha candidate); the 8 candidates of an instance sit in 8
//      adjacent lanes, the first-accept rule is a ballot, the winning trajectory is copied by the whole warp
//      (iLQR::forwardPassLineSearch, ilqr.cpp:311-361) ----
#ifndef H1_SEQ_THREADS
#define H1_SEQ_THREADS 128
#endif
constexpr int SEQ_THREADS = H1_SEQ_THREADS;
static_assert(H1ILQR_NALPHA == 8, "candidate groups are 8 lanes wide");
static_assert(NX % 3 == 0, "feedback loop is unrolled by 3");
#ifndef H1_SEQ_MINB
#define H1_SEQ_MINB 2
#endif
__global__ void __launch_bounds__(SEQ_THREADS, H1_SEQ_MINB)
k_line_search_seq(const DynModel* gmd, const H1Weights* gw, const H1SolverOptions* gopt, RefTable refs, int B, int N,
                  const int* __restrict__ mask, const int* __restrict__ list, const int* __restrict__ list_count,
                  const double* __restrict__ x0, const double* __restrict__ baseline,
                  double* __restrict__ xbar, double* __restrict__ ubar, const double* __restrict__ K,
                  const double* __restrict__ kff, double* __restrict__ xnew, double* __restrict__ unew,
                  int* __restrict__ ls_ok, double* __restrict__ ls_cost, int* __restrict__ ls_alpha) {
  extern __shared__ __align__(16) unsigned char smem[];
  // `list` (optional): compact list of the instances to search, built on the device by k_solve_state; slot s of
  // the grid then works on instance list[s] and full warps are formed however sparse the active set is
  const int nlist = list ? *list_count : B;
  if ((long)blockIdx.x * blockDim.x >= (long)nlist * H1ILQR_NALPHA) return;
  const DynModel* md;
  stage_model(smem, gmd, &md);
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int slot = (int)(g >> 3), cand = (int)(g & 7);
  const int inst = list ? (slot < nlist ? list[slot] : B) : slot;
  const bool act = inst < B && (list || !mask || mask[inst]);
  const int instc = min(inst, B - 1);
  double* xb = xbar + (size_t)instc * (N + 1) * NX;
  double* ub = ubar + (size_t)instc * N * NU;
  double* xn = xnew + ((size_t)instc * H1ILQR_NALPHA + cand) * (N + 1) * NX;
  double* un = unew + ((size_t)instc * H1ILQR_NALPHA + cand) * N * NU;
  double total = 0.0;
  // the thread's current state lives in shared memory (odd per-thread stride: conflict free); every f_D evaluation
  // and cost term reads it from there instead of going back to the global trajectory
  double* xs = reinterpret_cast<double*>(smem + ((sizeof(DynModel) + 15) / 16) * 16) + threadIdx.x * NX;
  // The 4 warps of a CTA walk the knots in step (a barrier per knot): f_D is ~20 k straight-line instructions, far more
  // than the instruction cache holds, and warps that sit at the same place of it share every fetched line (line search
  // 80 -> 77 ms per solve at 8192 instances; a second barrier inside the knot or 256-thread CTAs measured no better). CTAs
  // of 128 threads also leave whole SMs to the derivative kernels that run beside the second attempts.
  const double alpha = gopt->alphas[cand];
  const RefView r = refs.view(instc);
  if (act)
    for (int i = 0; i < NX; ++i) { const double v = x0 ? x0[(size_t)inst * NX + i] : xb[i]; xs[i] = v; xn[i] = v; }
  double u[NU], com[3];
#pragma unroll 1
  for (int t = 0; t < N; ++t) {
    if (act) {
      const double* Kt = K + ((size_t)inst * N + t) * NU * NX;
      const double* kt = kff + ((size_t)inst * N + t) * NU;
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = ub[t * NU + i] + alpha * kt[i];
      // u += K_t (x - xbar_t): three state entries per trip so that 57 loads of K are in flight per thread
#pragma unroll 1
      for (int l = 0; l < NX; l += 3) {
        const double dx0 = xs[l] - xb[t * NX + l];
        const double dx1 = xs[l + 1] - xb[t * NX + l + 1];
        const double dx2 = xs[l + 2] - xb[t * NX + l + 2];
        const double* Kl = Kt + l * NU;
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] += Kl[i] * dx0 + Kl[NU + i] * dx1 + Kl[2 * NU + i] * dx2;
      }
#pragma unroll
      for (int i = 0; i < NU; ++i) un[t * NU + i] = u[i];
    }
    if (act) {
      double* xnext = xn + (t + 1) * NX;
      dyn_step_seq(*md, xs, u, xnext, nullptr, com);
      total += knot_cost_seq(*md, *gw, r, t, xs, u, com, false);
      for (int i = 0; i < NX; ++i) xs[i] = xnext[i];
    }
#ifndef H1_SEQ_NOSYNC
    __syncthreads();
#endif
  }
  if (act) {
    dyn_com_seq(*md, xs, com);
    total += knot_cost_seq(*md, *gw, r, N, xs, nullptr, com, true);
  }
  __syncwarp();
  const double base = act ? baseline[inst] : 0.0;
  const bool better = act && (total < base - gopt->accept_margin);
  const unsigned votes = (__ballot_sync(0xffffffffu, better) >> (lane & ~7)) & 0xffu;
  const int win = votes ? (__ffs((int)votes) - 1) : -1;
  const double win_cost = __shfl_sync(0xffffffffu, total, (lane & ~7) + max(win, 0));
  if (act && cand == 0) {
    ls_ok[inst] = win >= 0;
    ls_cost[inst] = win >= 0 ? win_cost : base;
    ls_alpha[inst] = win;
  }
  // winners -> xbar / ubar, one 8-lane group (= instance) after another, copied by the whole warp
  for (int grp = 0; grp < 4; ++grp) {
    const int gw_ = __shfl_sync(0xffffffffu, win, grp * 8);
    const int gi = __shfl_sync(0xffffffffu, act ? inst : -1, grp * 8);
    if (gi < 0 || gw_ < 0) continue;
    const double* xw = xnew + ((size_t)gi * H1ILQR_NALPHA + gw_) * (N + 1) * NX;
    const double* uw = unew + ((size_t)gi * H1ILQR_NALPHA + gw_) * N * NU;
    double* xd = xbar + (size_t)gi * (N + 1) * NX;
    double* ud = ubar + (size_t)gi * N * NU;
    for (int i = lane; i < (N + 1) * NX; i += 32) xd[i] = xw[i];
    for (int i = lane; i < N * NU; i += 32) ud[i] = uw[i];
  }
}

}  // namespace h1
