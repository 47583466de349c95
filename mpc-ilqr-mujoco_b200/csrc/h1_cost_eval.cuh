// Warp-level evaluation of the trajectory cost used by the line search and the convergence test
// (reference: iLQR::computeTotalCost, /root/reference/src/ilqr/ilqr.cpp:363-518, and
// RobotUtils::constraintCost, /root/reference/src/common/robot_utils.cpp:615-672).
// Quirk Q1 is preserved: the CoM / CoM-velocity / foot position / foot velocity terms that the derivatives
// include are NOT part of this cost. The balance term uses the DYNAMICS-model CoM (a by-product of the f_D
// assembly of the same state, DynWarp::com) and the raw base linear velocity (quirk Q7).
#pragma once
#include "h1_dyn.cuh"

namespace h1 {

// Per-instance view of the reference window (device pointers, already offset to the instance).
struct RefView {
  const double* x_ref;    // [N+1][51]
  const double* u_ref;    // [N][19]
  const double* com_ref;  // [N+1][3]
  const double* ee_ref;   // [N+1][2][3]
  const int* stance;      // [N+1][2]
  const double* com_vel_ref;  // [N+1][3]
};

struct RefTable {
  const double *x_ref, *u_ref, *com_ref, *ee_ref, *com_vel_ref;
  const int* stance;
  int shared;  // 1: one window for all instances
  int N;
  __device__ __forceinline__ RefView view(int inst) const {
    const size_t i = shared ? 0 : (size_t)inst;
    RefView r;
    r.x_ref = x_ref + i * (size_t)(N + 1) * NX;
    r.u_ref = u_ref + i * (size_t)N * NU;
    r.com_ref = com_ref + i * (size_t)(N + 1) * 3;
    r.ee_ref = ee_ref + i * (size_t)(N + 1) * 6;
    r.stance = stance + i * (size_t)(N + 1) * 2;
    r.com_vel_ref = com_vel_ref + i * (size_t)(N + 1) * 3;
    return r;
  }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double limit_pen(double val, double lo, double hi, double wgt) {
  const double margin = 0.1 * (hi - lo), lo_s = lo + margin, hi_s = hi - margin;
  double c = 0.0;
  if (val > hi_s) { const double viol = val - hi_s; c += wgt * viol * viol; }
  if (val < lo_s) { const double viol = lo_s - val; c += wgt * viol * viol; }
  return c;
}

// support centre from horizon-local stance flags / foot targets (ilqr.cpp:403-437); false = aerial phase
__device__ __forceinline__ bool support_centre(const RefView& r, int t, double* ps) {
  const bool ls = r.stance[2 * t] == 1, rs = r.stance[2 * t + 1] == 1;
  const double* l = r.ee_ref + 6 * t;
  const double* rr = l + 3;
  if (ls && rs) { ps[0] = 0.5 * (l[0] + rr[0]); ps[1] = 0.5 * (l[1] + rr[1]); }
  else if (ls) { ps[0] = l[0]; ps[1] = l[1]; }
  else if (rs) { ps[0] = rr[0]; ps[1] = rr[1]; }
  else return false;
  return true;
}

// Cost of knot t for the state staged in w (w.q, w.v raw; w.com valid, i.e. after ph_rows) and control u
// (global/shared pointer; nullptr at the terminal knot = zero control). All lanes return the warp total.
__device__ __forceinline__ double knot_cost_warp(const DynModel& md, const DynWarp& w, const H1Weights& wt,
                                                 const RefView& r, int t, const double* u, bool terminal) {
  const int lane = threadIdx.x & 31;
  const double* Qd = terminal ? wt.Qfdiag : wt.Qdiag;
  double acc = 0.0;
  for (int i = lane; i < NX; i += 32) {
    const double xi = (i < NQ) ? w.q[i] : w.v[i - NQ];
    const double e = xi - r.x_ref[t * NX + i];
    acc += 0.5 * e * Qd[i] * e;
  }
  if (const double* qo = weights_offdiag(wt)) {   // off-diagonal parts of full Q / R / Qf (0.5 e'Qe, ilqr.cpp:372-373, 441)
    const double* Qo = qo + (terminal ? QOFF_QF : 0);
    for (int i = lane; i < NX; i += 32) {
      double sacc = 0.0;
      for (int j = 0; j < NX; ++j) sacc += Qo[j * NX + i] * (((j < NQ) ? w.q[j] : w.v[j - NQ]) - r.x_ref[t * NX + j]);
      acc += 0.5 * (((i < NQ) ? w.q[i] : w.v[i - NQ]) - r.x_ref[t * NX + i]) * sacc;
    }
    if (!terminal && lane < NU) {
      double sacc = 0.0;
      for (int j = 0; j < NU; ++j) sacc += qo[QOFF_R + j * NU + lane] * (u[j] - r.u_ref[t * NU + j]);
      acc += 0.5 * (u[lane] - r.u_ref[t * NU + lane]) * sacc;
    }
  }
  if (lane < NU) {
    const double ui = terminal ? 0.0 : u[lane];
    if (!terminal) { const double e = ui - r.u_ref[t * NU + lane]; acc += 0.5 * e * wt.Rdiag[lane] * e; }
    acc += limit_pen(ui, md.ctrl_lo[lane], md.ctrl_hi[lane], wt.w_control_limits);
    const double lo = md.jnt_lo[lane], hi = md.jnt_hi[lane];
    if (isfinite(lo) && isfinite(hi) && lo < hi) acc += limit_pen(w.q[7 + lane], lo, hi, wt.w_joint_limits);
  }
  if (lane == 0) {
    if (wt.w_upright > 0.0) {
      const double qw = w.q[3], qx = w.q[4], qy = w.q[5], qz = w.q[6];
      const double z0 = 2.0 * (qx * qz + qw * qy), z1 = 2.0 * (qy * qz - qw * qx);
      const double z2 = (1.0 - 2.0 * (qx * qx + qy * qy)) - 1.0;
      acc += 0.5 * wt.w_upright * (z0 * z0 + z1 * z1 + z2 * z2);
    }
    if (wt.w_balance > 0.0) {
      double ps[2];
      if (support_centre(r, t, ps)) {
        const double om = sqrt(w.com[2] / 9.81);
        const double r0 = w.com[0] + w.v[0] * om - ps[0], r1 = w.com[1] + w.v[1] * om - ps[1];
        acc += 0.5 * wt.w_balance * (r0 * r0 + r1 * r1);
      }
    }
  }
  return warp_sum(acc);
}

}  // namespace h1
