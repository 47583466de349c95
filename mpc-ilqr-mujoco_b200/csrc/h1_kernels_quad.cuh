// Batched line search on the quad-cooperative f_D (h1_dyn_quad.cuh): ONE WARP PER MPC INSTANCE, its 32 lanes =
// 8 alpha candidates x 4 kinematic chains. Replaces iLQR::forwardPassLineSearch (/root/reference/src/ilqr/ilqr.cpp:
// 311-361) for large batches; same contract as k_line_search / k_line_search_seq.
//   * feedback law u = ubar + alpha k + K (x - xbar) of the 8 candidates of an instance = a 19 x 51 x 8 contraction:
//     exactly the m8n8k4 shape of the fp64 tensor core, 39 DMMA per knot with the candidates as the N dimension
//     (lane l supplies K[8m + l/4][4s + l%4] and dx_{cand l/4}[4s + l%4] — the (candidate, chain) split of the lanes
//     is the fragment layout), instead of 969 FMAs per candidate;
//   * every candidate's state lives in shared memory (8 x 51 doubles per warp), the per-joint quantities of the
//     articulated-body recursion in a per-lane shared-memory column: no local memory, no spills;
//   * cost of a knot: every lane adds the terms of its own chain's coordinates (lane 0 also the base, upright and
//     capture-point terms), one quad-sum at the end of the rollout; first-accept rule = one ballot.
#pragma once
#include "h1_cost_eval.cuh"
#include "h1_dyn_quad.cuh"
#include "h1_kernels_dyn.cuh"

namespace h1 {

constexpr int Q4_XS = 52, Q4_US = 20;   // (instances per CTA = template parameter WARPS; one CTA per SM)
struct Q4WarpSmem {
  double xs[H1ILQR_NALPHA][Q4_XS];      // current state of the 8 candidates
  double us[H1ILQR_NALPHA][Q4_US];      // their controls at the current knot
  double st[Q4_STORE][32];              // per-lane store of dyn_step_quad
};

// cost terms of knot t that belong to lane g's coordinates (iLQR::computeTotalCost / RobotUtils::constraintCost, same
// terms and quirks as knot_cost_warp / knot_cost_seq); the four lanes' values add up to the knot cost
__device__ __forceinline__ double knot_cost_quad(const DynModel& md, const H1Weights& wt, const RefView& r, int t, int g,
                                                 const double* __restrict__ x, const double* __restrict__ u,
                                                 const double* com, bool terminal) {
  const double* Qd = terminal ? wt.Qfdiag : wt.Qdiag;
  const double* xr = r.x_ref + t * NX;
  double acc = 0.0;
  if (const double* qo = weights_offdiag(wt)) {   // off-diagonal parts of full Q / R / Qf: rows i = g (mod 4) on this lane
    const double* Qo = qo + (terminal ? QOFF_QF : 0);
#pragma unroll 1
    for (int i = g; i < NX; i += 4) {
      double sacc = 0.0;
#pragma unroll 1
      for (int j = 0; j < NX; ++j) sacc += Qo[j * NX + i] * (x[j] - xr[j]);
      acc += 0.5 * (x[i] - xr[i]) * sacc;
    }
    if (!terminal) {
#pragma unroll 1
      for (int i = g; i < NU; i += 4) {
        double sacc = 0.0;
#pragma unroll 1
        for (int j = 0; j < NU; ++j) sacc += qo[QOFF_R + j * NU + i] * (u[j] - r.u_ref[t * NU + j]);
        acc += 0.5 * (u[i] - r.u_ref[t * NU + i]) * sacc;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < Q4_CHAIN; ++i) {
    if (g == 3 && i == 0) continue;   // the torso is lane 2's
    const int b = q4_body(g, i), c = b - 1;
    const double eq = x[6 + b] - xr[6 + b], ev = x[NQ + 5 + b] - xr[NQ + 5 + b];
    acc += 0.5 * eq * Qd[6 + b] * eq;
    acc += 0.5 * ev * Qd[NQ + 5 + b] * ev;
    const double ui = terminal ? 0.0 : u[c];
    if (!terminal) { const double e = ui - r.u_ref[t * NU + c]; acc += 0.5 * e * wt.Rdiag[c] * e; }
    acc += limit_pen(ui, md.ctrl_lo[c], md.ctrl_hi[c], wt.w_control_limits);
    const double lo = md.jnt_lo[c], hi = md.jnt_hi[c];
    if (isfinite(lo) && isfinite(hi) && lo < hi) acc += limit_pen(x[6 + b], lo, hi, wt.w_joint_limits);
  }
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 7; ++i) { const double e = x[i] - xr[i]; acc += 0.5 * e * Qd[i] * e; }
#pragma unroll
    for (int i = NQ; i < NQ + 6; ++i) { const double e = x[i] - xr[i]; acc += 0.5 * e * Qd[i] * e; }
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
  }
  return acc;
}

// ---- nominal rollout on the quad-cooperative f_D: four lanes per instance, eight instances per warp. Same contract as
//      k_rollout / k_rollout_seq without the factor output (iLQR::forwardRolloutNominal, ilqr.cpp:119-124, and the baseline
//      computeTotalCost): xbar[t+1] = f_D(xbar[t], ubar[t]) for t >= t_begin (from x0 when given), cost of the whole
//      trajectory when cost_out != nullptr. One thread per instance (k_rollout_seq) left 128 warps of 25 dependent 20 k-
//      instruction evaluations on the GPU; here an instance is four lanes and an evaluation a quarter as long. ----
constexpr int RQ_WARPS = 2;   // 16 instances per CTA
struct RQWarpSmem {
  double xs[H1ILQR_NALPHA][Q4_XS];
  double us[H1ILQR_NALPHA][Q4_US];
  double st[Q4_STORE][32];
};
__global__ void __launch_bounds__(RQ_WARPS * 32)
k_rollout_quad(const DynModel* gmd, const H1Weights* gw, RefTable refs, int B, int N, int t_begin,
               const int* __restrict__ active, const double* __restrict__ x0, double* __restrict__ xbar,
               const double* __restrict__ ubar, double* __restrict__ cost_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 2, g = lane & 3;
  const int inst = (blockIdx.x * RQ_WARPS + warp) * 8 + q;
  if (inst >= B) return;
  if (active && !active[inst]) return;
  RQWarpSmem& ws = reinterpret_cast<RQWarpSmem*>(p)[warp];
  const QuadWarp cx(0xfu << (4 * q));     // quads of a warp finish independently (mask / batch edge)
  double* xs = ws.xs[q];
  double* us = ws.us[q];
  double* st = &ws.st[0][lane];
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  const double* ub = ubar + (size_t)inst * N * NU;
  const RefView r = refs.view(inst);
  for (int i = g; i < NX; i += 4) { const double v = x0 ? x0[(size_t)inst * NX + i] : xb[i]; xs[i] = v; if (x0) xb[i] = v; }
  double total = 0.0;
  cx.sync();
#pragma unroll 1
  for (int t = 0; t < N; ++t) {
    for (int i = g; i < NU; i += 4) us[i] = ub[t * NU + i];
    cx.sync();
    double com[3];
    if (t >= t_begin) {
      double qn[Q4_CHAIN], vn[Q4_CHAIN], bn[13];
      dyn_step_quad(*md, cx, g, xs, us, st, 32, qn, vn, bn, com);
      if (cost_out) total += knot_cost_quad(*md, *gw, r, t, g, xs, us, com, false);
      cx.sync();
#pragma unroll
      for (int i = 0; i < Q4_CHAIN; ++i) {
        if (g == 3 && i == 0) continue;
        const int b = q4_body(g, i);
        xs[6 + b] = qn[i]; xs[NQ + 5 + b] = vn[i];
      }
      if (g == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) xs[i] = bn[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[NQ + i] = bn[7 + i];
      }
      cx.sync();
      double* xnext = xb + (t + 1) * NX;
      for (int i = g; i < NX; i += 4) xnext[i] = xs[i];
    } else {
      if (cost_out) {
        dyn_com_quad(*md, cx, g, xs, com);
        total += knot_cost_quad(*md, *gw, r, t, g, xs, us, com, false);
      }
      cx.sync();
      const double* xnext = xb + (t + 1) * NX;
      for (int i = g; i < NX; i += 4) xs[i] = xnext[i];
      cx.sync();
    }
  }
  if (cost_out) {
    double com[3];
    dyn_com_quad(*md, cx, g, xs, com);
    total += knot_cost_quad(*md, *gw, r, N, g, xs, nullptr, com, true);
    total = quad_sum(cx, total);
    if (g == 0) cost_out[inst] = total;
  }
}

template <int Q4_WARPS>
__global__ void __launch_bounds__(Q4_WARPS * 32, 1)
k_line_search_quad(const DynModel* gmd, const H1Weights* gw, const H1SolverOptions* gopt, RefTable refs, int B, int N,
                   const int* __restrict__ mask, const int* __restrict__ list, const int* __restrict__ list_count,
                   const double* __restrict__ x0, const double* __restrict__ baseline,
                   double* __restrict__ xbar, double* __restrict__ ubar, const double* __restrict__ K,
                   const double* __restrict__ kff, double* __restrict__ xnew, double* __restrict__ unew,
                   int* __restrict__ ls_ok, double* __restrict__ ls_cost, int* __restrict__ ls_alpha) {
  extern __shared__ __align__(16) unsigned char smem[];
  // `list` (optional): compact list of the instances to search (k_solve_state); slot s works on instance list[s]
  const int nlist = list ? *list_count : B;
  if ((long)blockIdx.x * Q4_WARPS >= nlist) return;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * Q4_WARPS + warp;
  if (slot >= nlist) return;
  const int inst = list ? list[slot] : slot;
  if (!list && mask && !mask[inst]) return;
  Q4WarpSmem& ws = reinterpret_cast<Q4WarpSmem*>(p)[warp];
  const int cand = lane >> 2, g = lane & 3;
  const QuadWarp cx;
  double* xs = ws.xs[cand];
  double* us = ws.us[cand];
  double* st = &ws.st[0][lane];
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  double* ub = ubar + (size_t)inst * N * NU;
  double* xn = xnew + ((size_t)inst * H1ILQR_NALPHA + cand) * (N + 1) * NX;
  const RefView r = refs.view(inst);
  for (int i = g; i < NX; i += 4) { const double v = x0 ? x0[(size_t)inst * NX + i] : xb[i]; xs[i] = v; xn[i] = v; }
  const double al0 = gopt->alphas[2 * g], al1 = gopt->alphas[2 * g + 1];
  double total = 0.0;
  __syncwarp();
#pragma unroll 1
  for (int t = 0; t < N; ++t) {
    {  // u = ubar_t + alpha k_t + K_t (x - xbar_t) for the 8 candidates: [19 x 51] x [51 x 8] on the fp64 tensor core
      const double* Kt = K + ((size_t)inst * N + t) * NU * NX;
      const double* kt = kff + ((size_t)inst * N + t) * NU;
      const double* xbt = xb + t * NX;
      double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, c20 = 0.0, c21 = 0.0;
      const int r2 = 16 + cand;   // third row tile: controls 16..18 only
#pragma unroll
      for (int ks = 0; ks < (NX + 3) / 4; ++ks) {
        const int col = 4 * ks + g;
        const bool cok = col < NX;
        const double dx = cok ? xs[col] - xbt[col] : 0.0;
        const double* Kc = Kt + col * NU;
        const double a0 = cok ? Kc[cand] : 0.0;
        const double a1 = cok ? Kc[8 + cand] : 0.0;
        const double a2 = (cok && r2 < NU) ? Kc[r2] : 0.0;
        dmma884(c00, c01, a0, dx);
        dmma884(c10, c11, a1, dx);
        dmma884(c20, c21, a2, dx);
      }
      // this lane holds u[control 8m + cand][candidates 2g, 2g + 1]
      double* un0 = unew + ((size_t)inst * H1ILQR_NALPHA + 2 * g) * N * NU + t * NU;
      double* un1 = un0 + N * NU;
#define Q4_PUT(CTRL, C0, C1)                                                     \
  {                                                                              \
    const double ubv = ub[t * NU + (CTRL)], kv = kt[(CTRL)];                     \
    const double u0 = ubv + al0 * kv + (C0), u1 = ubv + al1 * kv + (C1);         \
    ws.us[2 * g][(CTRL)] = u0; ws.us[2 * g + 1][(CTRL)] = u1;                    \
    un0[(CTRL)] = u0; un1[(CTRL)] = u1;                                          \
  }
      Q4_PUT(cand, c00, c01)
      Q4_PUT(8 + cand, c10, c11)
      if (r2 < NU) Q4_PUT(r2, c20, c21)
#undef Q4_PUT
    }
    __syncwarp();
    double qn[Q4_CHAIN], vn[Q4_CHAIN], bn[13], com[3];
    dyn_step_quad(*md, cx, g, xs, us, st, 32, qn, vn, bn, com);
    total += knot_cost_quad(*md, *gw, r, t, g, xs, us, com, false);
    __syncwarp();   // every lane of the evaluation has finished reading the old state
#pragma unroll
    for (int i = 0; i < Q4_CHAIN; ++i) {
      if (g == 3 && i == 0) continue;
      const int b = q4_body(g, i);
      xs[6 + b] = qn[i]; xs[NQ + 5 + b] = vn[i];
    }
    if (g == 0) {
#pragma unroll
      for (int i = 0; i < 7; ++i) xs[i] = bn[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) xs[NQ + i] = bn[7 + i];
    }
    __syncwarp();
    double* xnext = xn + (t + 1) * NX;
    for (int i = g; i < NX; i += 4) xnext[i] = xs[i];
  }
  {
    double com[3];
    dyn_com_quad(*md, cx, g, xs, com);
    total += knot_cost_quad(*md, *gw, r, N, g, xs, nullptr, com, true);
  }
  total = quad_sum(cx, total);
  const double base = baseline[inst];
  const bool better = (g == 0) && (total < base - gopt->accept_margin);
  const unsigned votes = __ballot_sync(0xffffffffu, better);      // bit 4c = candidate c accepted
  const int win = votes ? ((__ffs((int)votes) - 1) >> 2) : -1;    // FIRST alpha in list order
  const double win_cost = __shfl_sync(0xffffffffu, total, 4 * max(win, 0));
  if (lane == 0) {
    ls_ok[inst] = win >= 0;
    ls_cost[inst] = win >= 0 ? win_cost : base;
    ls_alpha[inst] = win;
  }
  if (win < 0) return;
  __syncwarp();
  const double* xw = xnew + ((size_t)inst * H1ILQR_NALPHA + win) * (N + 1) * NX;
  const double* uw = unew + ((size_t)inst * H1ILQR_NALPHA + win) * N * NU;
  __threadfence_block();
  // the accepted candidate becomes the nominal trajectory: eight independent loads in flight per lane (a plain copy loop waits
  // for every load before it issues the next one: 57 dependent round trips, 5 % of this kernel's stall samples)
  auto copy8 = [&](double* __restrict__ dst, const double* __restrict__ src, int n) {
    for (int i0 = lane; i0 < n; i0 += 256) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (i0 + 32 * q < n) ? src[i0 + 32 * q] : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) if (i0 + 32 * q < n) dst[i0 + 32 * q] = v[q];
    }
  };
  copy8(xb, xw, (N + 1) * NX);
  copy8(ub, uw, N * NU);
}

}  // namespace h1
