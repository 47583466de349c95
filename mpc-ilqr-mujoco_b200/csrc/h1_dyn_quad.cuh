// Quad-cooperative evaluation of the one-step dynamics map f_D: FOUR lanes per evaluation, one per
// kinematic chain of H1 — the throughput form used by the batched line search (k_line_search_quad), where a warp
// carries the 8 alpha candidates of one MPC instance (8 x 4 lanes). Replaces RobotUtils::rolloutOneStep
// (/root/reference/src/common/robot_utils.cpp:106-117) inside iLQR::forwardPassLineSearch (ilqr.cpp:311-361).
//
// Why four lanes: one thread per evaluation (h1_dyn_seq.cuh) needs the state of all 20 bodies plus a 2.6 KB factor
// per thread — 255 registers and 4.4 KB of local memory, whose write-once / read-once traffic overflowed the L2
// (profiles/r01i: 4.2x the algorithmic DRAM bytes, 44 % long-scoreboard stalls). H1's tree is a base with four
// serial chains of five hinges each when the torso is counted with both arms:
//     lane 0: left leg  (bodies 1-5, foot 5)        lane 2: torso + left arm  (11, 12-15)
//     lane 1: right leg (bodies 6-10, foot 10)      lane 3: torso + right arm (11, 16-19)
// so every lane runs the SAME code on its own chain (no divergence except the sole contact of the two leg lanes),
// holds one chain's state in registers, and the lanes meet only twice: the two arm lanes add their articulated
// inertias at the shared torso (27 values, one xor-shuffle), and all four add their contributions to the 6 x 6 base
// block (30 values, two xor-shuffles).
//
// Algorithm per lane: articulated-body recursion in the common frame of h1_dyn_seq.cuh (origin = base origin,
// world-aligned axes, motion vectors [omega; v_O], forces [n_O; f]):
//   walk down the chain (poses, velocities, velocity-product accelerations, motion subspaces S_i);
//   walk back up: IA = I_i (+ K_f of the sole contact, folded into the foot exactly as in h1_dyn_seq.cuh) + Ia_child,
//     pA = f_i (- contact wrench) + pa_child;  U = IA S, D = S.U + armature + h d, u = tau - d v - S.pA,
//     Ia = IA - U U'/D, pa = pA + U u/D; the pose of the parent body is recovered by the INVERSE joint transform
//     (no per-body state is kept: 5 x 16 doubles of body inertias / wrenches would not fit the register file);
//   base: 6 x 6 system S_b' IA_0 S_b a_b = -S_b' pA_0 - d v (every lane solves it redundantly);
//   walk down again: a_i = (u_i - U_i.A_parent)/D_i, A_i = A_parent + S_i a_i;  semi-implicit Euler.
// This solves the same linearly-implicit system Mhat a = rhs as the L'DL factorisations of h1_dyn.cuh /
// h1_dyn_seq.cuh (Mhat = CRBA + armature + h D + sum_i J_i' W_i J_i); it never forms Mhat, so it produces no factor
// for the linearization (the line search needs none). Per joint it keeps S (6), U/D (6) and u/D (1) in a per-lane
// shared-memory column (65 doubles, stride 32 -> conflict free).
// Requires DynModel::seq_ok (H1's chain structure). The exchange primitives are a template parameter so that the
// same source runs under tests/emul with four host threads.
#pragma once
#include "h1_dyn_seq.cuh"

namespace h1 {

constexpr int Q4_CHAIN = 5;                       // hinges per lane
constexpr int Q4_STORE = 13 * Q4_CHAIN;           // per-lane store: S[5][6], U/D[5][6], u/D[5]
static_assert(SEQ_LEG_LEN == Q4_CHAIN && SEQ_ARM_LEN + 1 == Q4_CHAIN, "H1 chain structure");

// body of chain position i (0 = nearest to the base) of lane g
H1_HD int q4_body(int g, int i) { return g < 2 ? 1 + SEQ_LEG_LEN * g + i : (i == 0 ? SEQ_TORSO : SEQ_TORSO + SEQ_ARM_LEN * (g - 2) + i); }

#if defined(__CUDACC__)
struct QuadWarp {   // exchange between the 4 lanes of an evaluation (lanes 4c .. 4c+3 of a warp)
  unsigned mask;    // lanes that take part: the whole warp (line search: every quad is busy) or this quad only
  __device__ __forceinline__ explicit QuadWarp(unsigned m = 0xffffffffu) : mask(m) {}
  __device__ __forceinline__ double xor1(double v) const { return __shfl_xor_sync(mask, v, 1); }
  __device__ __forceinline__ double xor2(double v) const { return __shfl_xor_sync(mask, v, 2); }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
#endif

// sum over the four lanes of an evaluation; every lane obtains the same bits ((v0 + v1) + (v2 + v3), commutative pairs)
template <class CX> H1_DEV double quad_sum(const CX& cx, double v) {
  v += cx.xor1(v);
  v += cx.xor2(v);
  return v;
}

// packed symmetric 6 x 6 (lower triangle, row-major: K[i(i+1)/2 + j], j <= i) acting on motion vectors
H1_DEV void sym6_add_rigid(double* K, const double* I) {   // K += [[J, hx], [-hx, m 1]] of the rigid inertia I = (m, h, J)
  const double m = I[0], h0 = I[1], h1 = I[2], h2 = I[3];
  K[0] += I[4]; K[1] += I[7]; K[2] += I[5]; K[3] += I[8]; K[4] += I[9]; K[5] += I[6];
  K[7] += h2; K[8] -= h1; K[9] += m;
  K[10] -= h2; K[12] += h0; K[14] += m;
  K[15] += h1; K[16] -= h0; K[20] += m;
}
H1_DEV void sym6_apply(const double* K, const double* S, double* U) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) a += ((j <= i) ? K[i * (i + 1) / 2 + j] : K[j * (j + 1) / 2 + i]) * S[j];
    U[i] = a;
  }
}

// hinge of body b: parent state -> state of b (in place), motion subspace S; sin / cos are returned for the way back
H1_DEV void q4_joint(const DynModel& md, int b, const double* __restrict__ x, SeqBodyState& c, double* S, double* sn_out,
                     double* cs_out) {
  const double* p = md.pos[b];
  c.r[0] += c.R[0] * p[0] + c.R[1] * p[1] + c.R[2] * p[2];
  c.r[1] += c.R[3] * p[0] + c.R[4] * p[1] + c.R[5] * p[2];
  c.r[2] += c.R[6] * p[0] + c.R[7] * p[1] + c.R[8] * p[2];
  if (md.has_rfix[b]) {
    const double* Fx = md.rfix[b];
    double Tm[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        Tm[3 * i + k] = c.R[3 * i] * Fx[k] + c.R[3 * i + 1] * Fx[3 + k] + c.R[3 * i + 2] * Fx[6 + k];
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = Tm[i];
  }
  double sn, cs;
  sincos_t(x[6 + b], &sn, &cs);
  *sn_out = sn; *cs_out = cs;
  const int ax = md.axis[b];
  rot_right(c.R, ax, sn, cs);
  col_of(c.R, ax, S);
  cross_m(c.r, S, S + 3);
  const double vj = x[NQ + 5 + b];
#pragma unroll
  for (int i = 0; i < 6; ++i) c.V[i] += S[i] * vj;
  double c1[3], c2[3], c3[3];
  cross_m(c.V, S, c1); cross_m(c.V, S + 3, c2); cross_m(c.V + 3, S, c3);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c.Ab[i] += c1[i] * vj;
    c.Ab[3 + i] += (c2[i] + c3[i]) * vj;
  }
}
// inverse of q4_joint: state of body b -> state of its parent (same arithmetic mirrored; agrees with the way down to
// rounding, which is all the articulated-body recursion needs)
H1_DEV void q4_unjoint(const DynModel& md, int b, const double* __restrict__ x, SeqBodyState& c, const double* S, double sn,
                       double cs) {
  const double vj = x[NQ + 5 + b];
  double c1[3], c2[3], c3[3];
  cross_m(c.V, S, c1); cross_m(c.V, S + 3, c2); cross_m(c.V + 3, S, c3);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c.Ab[i] -= c1[i] * vj;
    c.Ab[3 + i] -= (c2[i] + c3[i]) * vj;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) c.V[i] -= S[i] * vj;
  rot_right(c.R, md.axis[b], -sn, cs);
  if (md.has_rfix[b]) {   // R_parent = R Fx'
    const double* Fx = md.rfix[b];
    double Tm[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        Tm[3 * i + k] = c.R[3 * i] * Fx[3 * k] + c.R[3 * i + 1] * Fx[3 * k + 1] + c.R[3 * i + 2] * Fx[3 * k + 2];
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = Tm[i];
  }
  const double* p = md.pos[b];
  c.r[0] -= c.R[0] * p[0] + c.R[1] * p[1] + c.R[2] * p[2];
  c.r[1] -= c.R[3] * p[0] + c.R[4] * p[1] + c.R[5] * p[2];
  c.r[2] -= c.R[6] * p[0] + c.R[7] * p[1] + c.R[8] * p[2];
}
// sole contact of foot f (its ankle body is in state c), folded into the foot body: pA -= [rho x phi; phi],
// IA += sum_k X_k' W_k X_k (h1_dyn_seq.cuh, seq_contact)
H1_DEV void q4_contact(const DynModel& md, int f, const SeqBodyState& c, double qz, double* pA, double* IA) {
  const double h = md.h;
#pragma unroll 1
  for (int k = 0; k < H1_NCP; ++k) {
    const double* pt = md.foot_pts[f * H1_NCP + k];
    const double rho[3] = {c.r[0] + c.R[0] * pt[0] + c.R[1] * pt[1] + c.R[2] * pt[2],
                           c.r[1] + c.R[3] * pt[0] + c.R[4] * pt[1] + c.R[5] * pt[2],
                           c.r[2] + c.R[6] * pt[0] + c.R[7] * pt[1] + c.R[8] * pt[2]};
    double t1[3];
    cross_m(c.V, rho, t1);
    const double pd[3] = {c.V[3] + t1[0], c.V[4] + t1[1], c.V[5] + t1[2]};
    const double dd_ = -(qz + rho[2]);
    const double s2 = dd_ * dd_ + md.eps * md.eps;
    double root, ratio;   // from ONE reciprocal square root (each within ~1.5 ulp) instead of an fp64 square root followed by a division
    root_and_ratio(s2, dd_, &root, &ratio);
    const double sp = 0.5 * (dd_ + root), al = 0.5 * (1.0 + ratio);
    const double W[3] = {al * (h * md.bt), al * (h * md.bt), al * (h * md.bn + h * h * md.kn)};
    const double phi[3] = {-(al * md.bt) * pd[0], -(al * md.bt) * pd[1], md.kn * sp - al * (md.bn + h * md.kn) * pd[2]};
    double n[3];
    cross_m(rho, phi, n);
#pragma unroll
    for (int i = 0; i < 3; ++i) { pA[i] -= n[i]; pA[3 + i] -= phi[i]; }
    contact_inertia_add(rho, W, IA);
  }
}

H1_DEV void q4_base_state(const DynModel& md, const double* __restrict__ x, SeqBodyState& base) {
  double qn[4];
  quat_normalize(x + 3, qn);
  quat_to_mat(qn, base.R);
  base.r[0] = base.r[1] = base.r[2] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    base.V[i] = base.R[3 * i] * x[NQ + 3] + base.R[3 * i + 1] * x[NQ + 4] + base.R[3 * i + 2] * x[NQ + 5];
    base.V[3 + i] = x[NQ + i];
  }
  double vxw[3];
  cross_m(base.V + 3, base.V, vxw);
#pragma unroll
  for (int i = 0; i < 3; ++i) { base.Ab[i] = 0.0; base.Ab[3 + i] = vxw[i] - md.gravity[i]; }
}

// x_next = f_D(x, u) for the evaluation this lane belongs to. g: lane within the evaluation (0..3); x: the 51 raw
// state entries and u: the 19 controls (nullptr = zero torques) of the evaluation, readable by its four lanes
// (shared memory); st: this lane's store column (Q4_STORE entries, stride qs doubles).
// Outputs: qn[i], vn[i] = next angle / rate of the lane's hinges (chain order; lane 3's entry 0 duplicates lane 2's
// torso), bn[13] = next base position (3), quaternion (4) and base velocity (6) — valid on lane 0 only —,
// com[3] = dynamics-model CoM of x (world), on every lane. Nothing is written to x: the caller stores the new state
// after its own synchronisation.
template <class CX>
H1_DEV void dyn_step_quad(const DynModel& md, const CX& cx, int g, const double* __restrict__ x, const double* __restrict__ u,
                          double* __restrict__ st, int qs, double* __restrict__ qn, double* __restrict__ vn,
                          double* __restrict__ bn, double* __restrict__ com) {
  const double h = md.h;
  const bool arm = g >= 2;
  SeqBodyState c;
  q4_base_state(md, x, c);
  double IA[21], pA[6], hs[3];
  hs[0] = hs[1] = hs[2] = 0.0;
  // ---- way down the chain ----
  double sn[Q4_CHAIN], cs[Q4_CHAIN];
#pragma unroll
  for (int i = 0; i < Q4_CHAIN; ++i) {
    double S[6];
    q4_joint(md, q4_body(g, i), x, c, S, &sn[i], &cs[i]);
#pragma unroll
    for (int k = 0; k < 6; ++k) st[(6 * i + k) * qs] = S[k];
  }
  // ---- way back up: articulated-body inertias / bias forces ----
#pragma unroll
  for (int i = 0; i < 21; ++i) IA[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) pA[i] = 0.0;
  // NOT unrolled: the five bodies of the chain share one copy of this (large) body. Unrolled, the line-search kernel needed 255
  // registers + 490 B of spills and 142 KB of code (9.6 % of its stall samples were instruction fetches); rolled: 248 registers,
  // no spills, 113 KB — 4.14 -> 3.68 ms per 8192-instance line search. (Rolling the way down as well: 3.76 ms.)
#pragma unroll 1
  for (int i = Q4_CHAIN - 1; i >= 0; --i) {
    const int b = q4_body(g, i), j = 5 + b;
    if (i == 0) {   // the torso carries both arms: the two arm lanes add what they bring up (leg lanes keep their own)
#pragma unroll
      for (int q = 0; q < 21; ++q) { const double t = cx.xor1(IA[q]); if (arm) IA[q] += t; }
#pragma unroll
      for (int q = 0; q < 6; ++q) { const double t = cx.xor1(pA[q]); if (arm) pA[q] += t; }
    }
    {
      double own[16];
      seq_body(md, b, c, own);
      sym6_add_rigid(IA, own + 6);
#pragma unroll
      for (int q = 0; q < 6; ++q) pA[q] += own[q];
      if (!(g == 3 && i == 0)) { hs[0] += own[7]; hs[1] += own[8]; hs[2] += own[9]; }   // the torso's mass counts once
    }
    if (i == Q4_CHAIN - 1 && !arm) q4_contact(md, g, c, x[2], pA, IA);
    double S[6], U[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) S[k] = st[(6 * i + k) * qs];
    sym6_apply(IA, S, U);
    const double D = S[0] * U[0] + S[1] * U[1] + S[2] * U[2] + S[3] * U[3] + S[4] * U[4] + S[5] * U[5] +
                     md.armature[j] + h * md.damping[j];
    const double uj = seq_tau(md, u, j) - md.damping[j] * x[NQ + j] -
                      (S[0] * pA[0] + S[1] * pA[1] + S[2] * pA[2] + S[3] * pA[3] + S[4] * pA[4] + S[5] * pA[5]);
    const double inv = 1.0 / D;
    double Ut[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { Ut[k] = U[k] * inv; st[(30 + 6 * i + k) * qs] = Ut[k]; }
    st[(60 + i) * qs] = uj * inv;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int q = 0; q <= r; ++q) IA[r * (r + 1) / 2 + q] -= Ut[r] * U[q];
#pragma unroll
    for (int k = 0; k < 6; ++k) pA[k] += Ut[k] * uj;
    if (i > 0) q4_unjoint(md, b, x, c, S, sn[i], cs[i]);
  }
  // ---- base block: lane 3 brings nothing (its torso result duplicates lane 2's), lane 0 adds the base body itself ----
  if (g == 3) {
#pragma unroll
    for (int q = 0; q < 21; ++q) IA[q] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) pA[q] = 0.0;
  }
  if (g == 0) {
    double own[16];
    q4_base_state(md, x, c);
    seq_body(md, 0, c, own);
    sym6_add_rigid(IA, own + 6);
#pragma unroll
    for (int q = 0; q < 6; ++q) pA[q] += own[q];
    hs[0] += own[7]; hs[1] += own[8]; hs[2] += own[9];
  }
#pragma unroll
  for (int q = 0; q < 21; ++q) IA[q] = quad_sum(cx, IA[q]);
#pragma unroll
  for (int q = 0; q < 6; ++q) pA[q] = quad_sum(cx, pA[q]);
#pragma unroll
  for (int q = 0; q < 3; ++q) hs[q] = quad_sum(cx, hs[q]);
  {
    const double inv = md.inv_total_mass;
    com[0] = x[0] + hs[0] * inv; com[1] = x[1] + hs[1] * inv; com[2] = x[2] + hs[2] * inv;
  }
  double ab[6], Ap[6];
  {
    double Rb[9];
    {
      double qq[4];
      quat_normalize(x + 3, qq);
      quat_to_mat(qq, Rb);
    }
    // M6 = S_b' IA_0 S_b (lower triangle), S_b = [0 e_s] (world-frame linear, s < 3), [R_b e_m; 0] (body-frame angular)
    double M[21], rhs[6];
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
      for (int s = 0; s <= jj; ++s) M[jj * (jj + 1) / 2 + s] = IA[(3 + jj) * (4 + jj) / 2 + 3 + s];
      rhs[jj] = -pA[3 + jj];
    }
    double KR[3][3];   // KR[a][m] = (Kaa R_b)[a][m]
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        double t = 0.0;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) t += ((bb <= a) ? IA[a * (a + 1) / 2 + bb] : IA[bb * (bb + 1) / 2 + a]) * Rb[3 * bb + m];
        KR[a][m] = t;
      }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int jj = 3 + m;
#pragma unroll
      for (int s = 0; s < 3; ++s)   // angular row, linear column: R_col_m . Kal[:, s], Kal[a][s] = Kla[s][a]
        M[jj * (jj + 1) / 2 + s] = Rb[m] * IA[(3 + s) * (4 + s) / 2] + Rb[3 + m] * IA[(3 + s) * (4 + s) / 2 + 1] +
                                   Rb[6 + m] * IA[(3 + s) * (4 + s) / 2 + 2];
#pragma unroll
      for (int s = 0; s <= m; ++s) M[jj * (jj + 1) / 2 + 3 + s] = Rb[m] * KR[0][s] + Rb[3 + m] * KR[1][s] + Rb[6 + m] * KR[2][s];
      rhs[jj] = -(Rb[m] * pA[0] + Rb[3 + m] * pA[1] + Rb[6 + m] * pA[2]);
    }
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) {
      M[jj * (jj + 1) / 2 + jj] += md.armature[jj] + h * md.damping[jj];
      rhs[jj] -= md.damping[jj] * x[NQ + jj];
    }
    // L'DL of the 6 x 6 block, last dof first (the order of h1_dyn_seq.cuh), fused forward substitution
#pragma unroll
    double dinv[6];   // reciprocals of the pivots: reused by the substitution below (fp64 divisions are ~20 instructions each)
#pragma unroll
    for (int k = 5; k >= 1; --k) {
      double* mk = M + k * (k + 1) / 2;
      const double inv = 1.0 / mk[k];
      dinv[k] = inv;
      double a[5];
#pragma unroll
      for (int s = 0; s < 5; ++s) if (s < k) a[s] = mk[s] * inv;
#pragma unroll
      for (int sj = 0; sj < 5; ++sj) {
        if (sj >= k) continue;
#pragma unroll
        for (int si = 0; si < 5; ++si) {
          if (si < sj || si >= k) continue;
          M[si * (si + 1) / 2 + sj] -= a[si] * mk[sj];
        }
      }
#pragma unroll
      for (int sj = 0; sj < 5; ++sj) {
        if (sj >= k) continue;
        mk[sj] = a[sj];
        rhs[sj] -= a[sj] * rhs[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double* mk = M + k * (k + 1) / 2;
      double a = rhs[k] * (k == 0 ? 1.0 / mk[0] : dinv[k]);
#pragma unroll
      for (int s = 0; s < k; ++s) a -= mk[s] * ab[s];
      ab[k] = a;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      Ap[a] = Rb[3 * a] * ab[3] + Rb[3 * a + 1] * ab[4] + Rb[3 * a + 2] * ab[5];
      Ap[3 + a] = ab[a];
    }
  }
  // ---- way down again: joint accelerations, semi-implicit Euler ----
#pragma unroll
  for (int i = 0; i < Q4_CHAIN; ++i) {
    const int b = q4_body(g, i);
    double a = st[(60 + i) * qs];
#pragma unroll
    for (int k = 0; k < 6; ++k) a -= st[(30 + 6 * i + k) * qs] * Ap[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) Ap[k] += st[(6 * i + k) * qs] * a;
    const double v1 = x[NQ + 5 + b] + h * a;
    vn[i] = v1;
    qn[i] = x[6 + b] + h * v1;
  }
  if (g == 0) {
    double wn[3];
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) {
      const double v1 = x[NQ + jj] + h * ab[jj];
      bn[7 + jj] = v1;
      if (jj < 3) bn[jj] = x[jj] + h * v1; else wn[jj - 3] = v1;
    }
    quat_step(x + 3, wn, h, bn + 3);
  }
}

// dynamics-model CoM only (terminal knot of the line-search cost), four lanes per state
template <class CX>
H1_DEV void dyn_com_quad(const DynModel& md, const CX& cx, int g, const double* __restrict__ x, double* __restrict__ com) {
  double R[9], r[3] = {0.0, 0.0, 0.0}, hs[3] = {0.0, 0.0, 0.0};
  {
    double qq[4];
    quat_normalize(x + 3, qq);
    quat_to_mat(qq, R);
  }
  if (g == 0) {
    const double* ip = md.ipos[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) hs[i] = md.mass[0] * (R[3 * i] * ip[0] + R[3 * i + 1] * ip[1] + R[3 * i + 2] * ip[2]);
  }
#pragma unroll 1
  for (int i = 0; i < Q4_CHAIN; ++i) {
    const int b = q4_body(g, i);
    const double* p = md.pos[b];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] += R[3 * k] * p[0] + R[3 * k + 1] * p[1] + R[3 * k + 2] * p[2];
    if (md.has_rfix[b]) {
      const double* Fx = md.rfix[b];
      double Tm[9];
      for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 3; ++k) Tm[3 * a + k] = R[3 * a] * Fx[k] + R[3 * a + 1] * Fx[3 + k] + R[3 * a + 2] * Fx[6 + k];
      for (int k = 0; k < 9; ++k) R[k] = Tm[k];
    }
    double sn, cs;
    sincos_t(x[6 + b], &sn, &cs);
    rot_right(R, md.axis[b], sn, cs);
    if (!(g == 3 && i == 0)) {
      const double* ip = md.ipos[b];
      const double m = md.mass[b];
#pragma unroll
      for (int k = 0; k < 3; ++k) hs[k] += m * (r[k] + R[3 * k] * ip[0] + R[3 * k + 1] * ip[1] + R[3 * k + 2] * ip[2]);
    }
  }
  const double inv = md.inv_total_mass;
#pragma unroll
  for (int k = 0; k < 3; ++k) com[k] = x[k] + quad_sum(cx, hs[k]) * inv;
}

}  // namespace h1
