// Cost-quadratics launch kernel, alpha-parallel line search, warm-start shift, per-instance solver state
// machine and first-control kernels.
#pragma once
#include "h1_costq.cuh"
#include "h1_kernels_dyn.cuh"

namespace h1 {

// ---- cost quadratics: two warps per (instance, knot) incl. the terminal knot; two knots per CTA ----
constexpr int CQ_WARPS = 4;
constexpr int CQ_KNOTS = CQ_WARPS * 32 / CQ_LANES;
#ifndef H1_CQ_MINB
#define H1_CQ_MINB 4   // 16 resident warps at 128 registers (r02t: 4.77 ms per 8192-instance launch; 3 CTAs / 165 registers: 5.42 ms)
#endif
template <bool FULLQ>
__global__ void __launch_bounds__(CQ_WARPS * 32, H1_CQ_MINB)
k_cost_quadratics(const CostModel* gcm, const DynModel* gmd, const H1Weights* gw, RefTable refs, int B, int N,
                  const int* __restrict__ active, const double* __restrict__ xbar, const double* __restrict__ ubar,
                  double* __restrict__ lx, double* __restrict__ lu, double* __restrict__ lxx,
                  double* __restrict__ luu, const double* __restrict__ qoff) {
  extern __shared__ __align__(16) unsigned char smem[];
  const CostModel* cm;
  unsigned char* p = stage_model(smem, gcm, &cm);
  CostWarp& w = reinterpret_cast<CostWarp*>(p)[threadIdx.x / CQ_LANES];
  const long gwarp = (long)blockIdx.x * CQ_KNOTS + (threadIdx.x / CQ_LANES);   // knot slot (both warps of a pair leave together)
  if (gwarp >= (long)B * (N + 1)) return;
  const int inst = (int)(gwarp / (N + 1)), t = (int)(gwarp % (N + 1));
  if (active && !active[inst]) return;
  const RefView r = refs.view(inst);
  const bool terminal = (t == N);
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const double* u = terminal ? nullptr : ubar + ((size_t)inst * N + t) * NU;
  KnotTargets kt;
  kt.com_ref = r.com_ref + 3 * t; kt.com_vel_ref = r.com_vel_ref + 3 * t; kt.ee_ref = r.ee_ref + 6 * t;
  kt.stance = r.stance + 2 * t; kt.terminal = terminal;
  cost_quadratics_warp<FULLQ>(*cm, *gmd, *gw, w, x, u, r.x_ref + t * NX, terminal ? nullptr : r.u_ref + t * NU, kt,
                       lx + ((size_t)inst * (N + 1) + t) * NX, terminal ? nullptr : lu + ((size_t)inst * N + t) * NU,
                       lxx + ((size_t)inst * (N + 1) + t) * LXX_STRIDE,
                       terminal ? nullptr : luu + ((size_t)inst * N + t) * NU * NU, qoff);
}

// ---- line search: one CTA per instance, one warp per alpha candidate, every candidate rolled out
//      concurrently with its cost reduced on the fly; the FIRST alpha (list order) with
//      cost < baseline - margin wins (iLQR::forwardPassLineSearch, ilqr.cpp:311-361). ----
__global__ void __launch_bounds__(H1ILQR_NALPHA * 32)
k_line_search(const DynModel* gmd, const H1Weights* gw, const H1SolverOptions* gopt, RefTable refs, int N,
              const int* __restrict__ mask, const double* __restrict__ x0, const double* __restrict__ baseline,
              double* __restrict__ xbar,
              double* __restrict__ ubar, const double* __restrict__ K, const double* __restrict__ kff,
              double* __restrict__ xnew, double* __restrict__ unew, int* __restrict__ ls_ok,
              double* __restrict__ ls_cost, int* __restrict__ ls_alpha) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int inst = blockIdx.x;
  if (mask && !mask[inst]) return;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[warp];
  double* shc = reinterpret_cast<double*>(p + sizeof(DynWarp) * H1ILQR_NALPHA);  // [NALPHA] costs
  double* dxs = shc + H1ILQR_NALPHA + warp * (NX + NU);                          // per-warp dx[51], u[19]
  __shared__ int winner;
  const double alpha = gopt->alphas[warp];
  const RefView r = refs.view(inst);
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  double* ub = ubar + (size_t)inst * N * NU;
  double* xn = xnew + ((size_t)inst * H1ILQR_NALPHA + warp) * (N + 1) * NX;
  double* un = unew + ((size_t)inst * H1ILQR_NALPHA + warp) * N * NU;
  for (int i = lane; i < NX; i += 32) xn[i] = x0 ? x0[(size_t)inst * NX + i] : xb[i];
  __syncwarp();
  double total = 0.0;
  for (int t = 0; t < N; ++t) {
    const double* Kt = K + ((size_t)inst * N + t) * NU * NX;
    for (int i = lane; i < NX; i += 32) dxs[i] = xn[t * NX + i] - xb[t * NX + i];
    __syncwarp();
    if (lane < NU) {
      double acc = 0.0;
      for (int l = 0; l < NX; ++l) acc += Kt[l * NU + lane] * dxs[l];
      const double uv = ub[t * NU + lane] + alpha * kff[((size_t)inst * N + t) * NU + lane] + acc;
      dxs[NX + lane] = uv;
      un[t * NU + lane] = uv;
    }
    __syncwarp();
    dyn_step_warp(*md, w, xn + t * NX, dxs + NX, xn + (t + 1) * NX);
    total += knot_cost_warp(*md, w, *gw, r, t, dxs + NX, false);
    __syncwarp();
  }
  dyn_assemble_warp(*md, w, xn + N * NX, nullptr);
  total += knot_cost_warp(*md, w, *gw, r, N, nullptr, true);
  if (lane == 0) shc[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double base = baseline[inst];
    int win = -1;
    for (int a = 0; a < H1ILQR_NALPHA; ++a)
      if (shc[a] < base - gopt->accept_margin) { win = a; break; }
    winner = win;
    ls_ok[inst] = win >= 0;
    ls_cost[inst] = win >= 0 ? shc[win] : base;
    ls_alpha[inst] = win;
  }
  __syncthreads();
  const int win = winner;
  if (win < 0) return;
  const double* xw = xnew + ((size_t)inst * H1ILQR_NALPHA + win) * (N + 1) * NX;
  const double* uw = unew + ((size_t)inst * H1ILQR_NALPHA + win) * N * NU;
  __threadfence_block();
  for (int i = threadIdx.x; i < (N + 1) * NX; i += blockDim.x) xb[i] = xw[i];
  for (int i = threadIdx.x; i < N * NU; i += blockDim.x) ub[i] = uw[i];
}

// ---- initial guess (iLQR::initializeWithReference, ilqr.cpp:50-117): warm = shift by one knot, cold =
//      constant control guess. The rollouts that complete the guess are k_rollout launches. ----
__global__ void k_init_guess(int B, int N, const double* __restrict__ x0, const int* __restrict__ warm,
                             const int* __restrict__ has_prev, const double* __restrict__ u_init, int u_shared,
                             const double* __restrict__ prev_xbar, const double* __restrict__ prev_ubar,
                             double* __restrict__ xbar, double* __restrict__ ubar, int* __restrict__ warm_mask,
                             int* __restrict__ cold_mask) {
  const int inst = blockIdx.x;
  const bool wm = (warm == nullptr || warm[inst]) && has_prev[inst];
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  double* ub = ubar + (size_t)inst * N * NU;
  const double* px = prev_xbar + (size_t)inst * (N + 1) * NX;
  const double* pu = prev_ubar + (size_t)inst * N * NU;
  for (int i = threadIdx.x; i < NX; i += blockDim.x) xb[i] = x0[(size_t)inst * NX + i];
  if (wm) {
    for (int i = threadIdx.x; i < N * NU; i += blockDim.x) {
      const int t = i / NU, c = i % NU;
      ub[i] = (t < N - 1) ? pu[(t + 1) * NU + c] : pu[(N - 1) * NU + c];
    }
    for (int i = threadIdx.x; i < (N - 1) * NX; i += blockDim.x) xb[NX + i] = px[2 * NX + i];
  } else {
    const double* ui = u_init + (u_shared ? 0 : (size_t)inst * NU);
    for (int i = threadIdx.x; i < N * NU; i += blockDim.x) ub[i] = ui[i % NU];
  }
  if (threadIdx.x == 0) { warm_mask[inst] = wm ? 1 : 0; cold_mask[inst] = wm ? 0 : 1; }
}

// ---- per-instance solver state machine of iLQR::solve (ilqr.cpp:547-656), one thread per instance ----
// The launch sequence is pipelined (h1ilqr_capi.cu, enqueue_solve): instances whose FIRST line search succeeded finish
// their iteration in phase 1 and their next linearization / cost quadratics start at once (`early` set), while the
// second attempts (lambda x 10, backward pass, line search) of the others run concurrently on a second stream and
// finish in phase 2; those that succeed there join late (`late` set). Instances whose two attempts both failed keep
// their trajectory, so their linearization and cost quadratics of this iteration stay valid (quirk Q10: `continue`).
struct SolveState {
  double* lambda; double* cost; double* prev_cost; double* nominal_cost;
  int* active; int* second; int* early; int* late; int* iters; int* status;
  int* ls_ok; double* ls_cost; int* ls_alpha;
  double* cost_trace; int* alpha_trace;
  int* act_list; int* sec_list; int* early_list; int* late_list;
  int* list_count;   // [4]: active, second, early, late — compact instance lists so that warps / CTAs are full however sparse a set is
};
// end of an iteration whose line search succeeded (ilqr.cpp:628-645); returns whether the instance stays active
__device__ __forceinline__ bool solve_state_accept(const SolveState& st, const H1SolverOptions& o, int i, int it) {
  const int maxit = o.max_iterations;
  const double cur = st.ls_cost[i];
  st.iters[i] = it + 1;
  st.cost[i] = cur;
  st.nominal_cost[i] = cur;   // baseline of the next line search = cost of the trajectory this accept installed
  st.lambda[i] = fmax(st.lambda[i] / 2.0, o.reg_min);
  st.cost_trace[(size_t)i * maxit + it] = cur;
  bool on = true;
  if (!isfinite(cur)) { st.status[i] = 1; on = false; }
  else if (fabs(cur - st.prev_cost[i]) < o.tolerance) on = false;
  else if (cur > o.divergence_cost) on = false;
  if (!on) st.active[i] = 0;
  return on;
}
// phase 0: iteration begin; 1: after the first line search; 2: after the second attempts
__global__ void k_solve_state(SolveState st, const H1SolverOptions* gopt, int B, int it, int phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const H1SolverOptions& o = *gopt;
  const int maxit = o.max_iterations;
  if (phase == 0) {
    if (it == 0) {
      st.active[i] = 1; st.iters[i] = 0; st.status[i] = 0;
      for (int k = 0; k < maxit; ++k) { st.cost_trace[(size_t)i * maxit + k] = 0.0; st.alpha_trace[((size_t)i * maxit + k) * 2] = -2; st.alpha_trace[((size_t)i * maxit + k) * 2 + 1] = -2; }
    }
    st.second[i] = 0;
    if (st.active[i]) {
      st.prev_cost[i] = st.cost[i];
      // The baseline of the line search (ilqr.cpp:318: computeTotalCost of the re-rolled-out trajectory) lives in
      // nominal_cost: the iteration-0 rollout from x0 writes it and every accepted candidate replaces it
      // (solve_state_accept). It is NOT cost[i] while no candidate has been accepted: cost[i] is then still the
      // cost of the initial guess BEFORE the rollout, which differs for a warm start whose measured state left
      // the predicted one (the shifted trajectory is not dynamics-consistent).
      st.act_list[atomicAdd(&st.list_count[0], 1)] = i;   // (the counter is zeroed before the phase-0 launch)
    }
    return;
  }
  if (phase == 1) {
    st.early[i] = 0; st.late[i] = 0;
    if (!st.active[i]) return;
    st.alpha_trace[((size_t)i * maxit + it) * 2] = st.ls_alpha[i];
    if (!st.ls_ok[i]) {
      st.lambda[i] = fmin(st.lambda[i] * 10.0, o.reg_max);
      st.second[i] = 1;
      st.sec_list[atomicAdd(&st.list_count[1], 1)] = i;
    } else if (solve_state_accept(st, o, i, it)) {
      st.early[i] = 1;
      st.early_list[atomicAdd(&st.list_count[2], 1)] = i;
    }
    return;
  }
  // phase 2: only the instances of the second attempt
  if (!st.active[i] || !st.second[i]) return;
  st.alpha_trace[((size_t)i * maxit + it) * 2 + 1] = st.ls_alpha[i];
  if (!st.ls_ok[i]) {  // both attempts failed
    st.iters[i] = it + 1;
    st.cost_trace[(size_t)i * maxit + it] = st.cost[i];
    if (it > 1) st.active[i] = 0;
    return;              // `continue`: no lambda decrease, no convergence test (quirk Q10)
  }
  if (solve_state_accept(st, o, i, it)) {
    st.late[i] = 1;
    st.late_list[atomicAdd(&st.list_count[3], 1)] = i;
  }
}

// ---- u_apply = ubar[0] + K[0] (x_measured - xbar[0]) and bookkeeping of the previous solution
//      (MPC::stepOnce, mpc.cpp:97-111) ----
__global__ void k_first_control(int B, int N, const double* __restrict__ x_meas, const double* __restrict__ xbar,
                                const double* __restrict__ ubar, const double* __restrict__ K,
                                double* __restrict__ u_apply) {
  const int inst = blockIdx.x, i = threadIdx.x;
  if (i >= NU) return;
  const double* xb = xbar + (size_t)inst * (N + 1) * NX;
  const double* K0 = K + (size_t)inst * N * NU * NX;
  double acc = 0.0;
  for (int l = 0; l < NX; ++l) acc += K0[l * NU + i] * (x_meas[(size_t)inst * NX + l] - xb[l]);
  u_apply[(size_t)inst * NU + i] = ubar[(size_t)inst * N * NU + i] + acc;
}

// ---- device-resident closed loop (main/humanoid_mpc.cpp:130-179 without host round trips) ----
// Full reference tables of RobotUtils (x_ref_full_, com_ref_full_, ee_pos_ref_full_, com_vel_ref_full_, contact_schedule_),
// uploaded once; every instance has its own time index.
struct RefTables {
  const double *x, *com, *ee, *cv;   // [rows][51], [rows][3], [rows][2][3], [rows][3]
  const int* contact;                // [contact_rows][2]
  int rows, contact_rows;
  int schedule_offset;               // 0: horizon-local schedule / foot / CoM-velocity lookups as the reference does (quirk Q6)
};
// Reference window of every instance at its time index: MPC::extractReferenceWindow / RobotUtils::getReferenceWindow
// (mpc.cpp:163-166, robot_utils.cpp:422-443: rows min(t_idx + i, last)), isStance / getEEReference / getCoMVelReference
// with the HORIZON-LOCAL knot index i (robot_utils.cpp:494-549). One thread per (instance, knot).
__global__ void k_extract_window(RefTables tb, int B, int N, const int* __restrict__ t_idx, double* __restrict__ x_ref,
                                 double* __restrict__ com_ref, double* __restrict__ ee_ref, double* __restrict__ cv_ref,
                                 int* __restrict__ stance) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)B * (N + 1)) return;
  const int inst = (int)(g / (N + 1)), i = (int)(g - (long)inst * (N + 1));
  const int row = min(t_idx[inst] + i, tb.rows - 1);
  const int loc = tb.schedule_offset ? row : min(i, tb.rows - 1);
  const size_t o = (size_t)inst * (N + 1) + i;
  for (int k = 0; k < NX; ++k) x_ref[o * NX + k] = tb.x[(size_t)row * NX + k];
  for (int k = 0; k < 3; ++k) { com_ref[o * 3 + k] = tb.com[(size_t)row * 3 + k]; cv_ref[o * 3 + k] = tb.cv[(size_t)loc * 3 + k]; }
  for (int k = 0; k < 6; ++k) ee_ref[o * 6 + k] = tb.ee[(size_t)loc * 6 + k];
  const int srow = tb.schedule_offset ? t_idx[inst] + i : i;   // isStance: rows past the schedule count as stance
  for (int e = 0; e < 2; ++e) stance[o * 2 + e] = (srow < tb.contact_rows) ? (tb.contact[srow * 2 + e] == 1) : 1;
}
// end of a closed-loop step: logs, t_idx_++ (mpc.cpp:113), step counter
__global__ void k_closed_loop_advance(int B, int* __restrict__ t_idx, int* __restrict__ step_ctr, const double* __restrict__ cost,
                                      const int* __restrict__ iters, const double* __restrict__ u_apply,
                                      double* __restrict__ log_cost, int* __restrict__ log_iters, double* __restrict__ log_u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int step = *step_ctr;
  if (i < B) {
    t_idx[i] += 1;
    if (log_cost) log_cost[(size_t)step * B + i] = cost[i];
    if (log_iters) log_iters[(size_t)step * B + i] = iters[i];
    if (log_u) for (int k = 0; k < NU; ++k) log_u[((size_t)step * B + i) * NU + k] = u_apply[(size_t)i * NU + k];
  }
  __syncthreads();
}
__global__ void k_increment(int* p) { *p += 1; }

// x0[i] = xbar[i][0] (stage timing: candidates of the line search start from the trajectory's own first state)
__global__ void k_copy_x0(int B, int N, const double* __restrict__ xbar, double* __restrict__ x0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  for (int k = 0; k < NX; ++k) x0[(size_t)i * NX + k] = xbar[(size_t)i * (N + 1) * NX + k];
}
// lxx is kept as its lower triangle on the device (k_cost_quadratics): fill the upper one for a caller that wants the matrix
__global__ void k_mirror_lower(long nmat, double* __restrict__ lxx) {
  const long m = blockIdx.x;
  if (m >= nmat) return;
  double* M = lxx + m * LXX_STRIDE;
  for (int i = threadIdx.x; i < NX * NX; i += blockDim.x) {
    const int c = i / NX, r = i - c * NX;
    if (r < c) M[i] = M[r * NX + c];
  }
}
__global__ void k_fill_int(int n, int* p, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_double(int n, double* p, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace h1
