// Thread-sequential evaluation of the one-step dynamics map f_D — the throughput form of
// RobotUtils::rolloutOneStep (/root/reference/src/common/robot_utils.cpp:106-117) for batched solves.
//
// The warp-cooperative form (h1_dyn.cuh) spends one warp per evaluation and keeps ~10 of its 32 lanes busy; it is
// the latency form. Here ONE THREAD evaluates one f_D and walks the kinematic tree once, in DFS order, so a warp
// carries 32 independent evaluations (candidate rollouts of the line search, instances of the nominal rollout)
// with no divergence, no shuffles and no shared-memory exchange:
//   * body poses, spatial velocities, bias accelerations, inertias and bias wrenches about the base origin in
//     world-aligned axes (the same quantities as ph_walk), so subtree totals are plain sums;
//   * the sole contact of a foot is folded into that foot body: its linearly-implicit force law
//     F_i = phi_i - W_i (J_i a) adds the wrench -[rho x phi; phi] to the body's bias wrench and the spatial
//     "inertia" K_f = sum_i X_i' W_i X_i (X_i = [rho_i x e_c; e_c]) to every composite inertia that contains the
//     foot, which yields Mhat = M + h D + sum_i J_i' W_i J_i and rhs = tau - c - D v + sum_i J_i' phi_i exactly;
//   * the walk is specialised for H1's chain structure (two 5-hinge legs and a torso below the base, two 4-hinge
//     arms below the torso; DynModel::seq_ok): a chain is walked down and back up with every index static, the
//     rows of Mhat of its dofs are produced on the way up and the chain is ELIMINATED at once (its steps of the
//     L'DL factorisation, forward substitution fused) while its rows, the rows of its ancestors (base 6 x 6,
//     torso) and the right-hand sides sit in registers; only L, D and rhs of the finished chain go to a
//     per-thread store for the back substitution;
//   * base 6 x 6 block last, back substitution root first, semi-implicit Euler + quaternion exponential.
// Optionally writes the factor (PrimalFactor) consumed by the linearization kernels and the dynamics-model CoM
// used by the line-search cost. No CUDA intrinsics: compiles as plain C++ for tests/emul.
#pragma once
#include "h1_lin_dirs.cuh"

namespace h1 {

struct SeqBodyState {
  double R[9], r[3], V[6], Ab[6];
};

// K += w_c X_c X_c' for the three world axes c, X_c = [rho x e_c; e_c]; K is the lower triangle, row-major
// (K[i(i+1)/2 + j], j <= i) of a symmetric 6x6 acting on motion vectors [omega; v].
H1_DEV void contact_inertia_add(const double* rho, const double* W, double* K) {
  const double X[3][3] = {{0.0, rho[2], -rho[1]}, {-rho[2], 0.0, rho[0]}, {rho[1], -rho[0], 0.0}};  // rho x e_c
  // angular-angular block
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= i; ++j)
      K[i * (i + 1) / 2 + j] += W[0] * X[0][i] * X[0][j] + W[1] * X[1][i] * X[1][j] + W[2] * X[2][i] * X[2][j];
  // linear-angular block: K[3 + c][j] = W_c X_c[j]
  for (int c = 0; c < 3; ++c)
    for (int j = 0; j < 3; ++j) K[(3 + c) * (4 + c) / 2 + j] += W[c] * X[c][j];
  // linear-linear block: diag(W)
  for (int c = 0; c < 3; ++c) K[(3 + c) * (4 + c) / 2 + 3 + c] += W[c];
}
// P += K S for the packed symmetric K
H1_DEV void contact_inertia_apply(const double* K, const double* S, double* P) {
  for (int i = 0; i < 6; ++i) {
    double a = 0.0;
    for (int j = 0; j < 6; ++j) a += ((j <= i) ? K[i * (i + 1) / 2 + j] : K[j * (j + 1) / 2 + i]) * S[j];
    P[i] += a;
  }
}

// ---- building blocks of the chain-structured walk ----
// hinge of body b: parent state -> state of b (in place) and its motion subspace S (about the base origin)
H1_DEV void seq_joint(const DynModel& md, int b, const double* __restrict__ x, SeqBodyState& c, double* S) {
  const double* p = md.pos[b];
  c.r[0] += c.R[0] * p[0] + c.R[1] * p[1] + c.R[2] * p[2];
  c.r[1] += c.R[3] * p[0] + c.R[4] * p[1] + c.R[5] * p[2];
  c.r[2] += c.R[6] * p[0] + c.R[7] * p[1] + c.R[8] * p[2];
  if (md.has_rfix[b]) {
    const double* Fx = md.rfix[b];
    double Tm[9];
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k)
        Tm[3 * i + k] = c.R[3 * i] * Fx[k] + c.R[3 * i + 1] * Fx[3 + k] + c.R[3 * i + 2] * Fx[6 + k];
    for (int i = 0; i < 9; ++i) c.R[i] = Tm[i];
  }
  double sn, cs;
  sincos_t(x[6 + b], &sn, &cs);
  const int ax = md.axis[b];
  rot_right(c.R, ax, sn, cs);
  col_of(c.R, ax, S);
  cross_m(c.r, S, S + 3);
  const double vj = x[NQ + 5 + b];
  for (int i = 0; i < 6; ++i) c.V[i] += S[i] * vj;
  double c1[3], c2[3], c3[3];
  cross_m(c.V, S, c1); cross_m(c.V, S + 3, c2); cross_m(c.V + 3, S, c3);
  for (int i = 0; i < 3; ++i) {
    c.Ab[i] += c1[i] * vj;
    c.Ab[3 + i] += (c2[i] + c3[i]) * vj;
  }
}
// bias wrench (6) and spatial inertia (10) of body b about the base origin
H1_DEV void seq_body(const DynModel& md, int b, const SeqBodyState& c, double* o) {
  body_inertia_seq(md, b, c.R, c.r, o + 6);
  body_wrench_seq(o + 6, c.V, c.Ab, o);
}
// sole contact of foot f (its ankle body is in state c): bias wrench -= [rho x phi; phi], Kf = sum X' W X
H1_DEV void seq_contact(const DynModel& md, int f, const SeqBodyState& c, double qz, double* o, double* Kf) {
  const double h = md.h;
  for (int i = 0; i < 21; ++i) Kf[i] = 0.0;
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
    double root, ratio;
    root_and_ratio(dd_ * dd_ + md.eps * md.eps, dd_, &root, &ratio);
    const double sp = 0.5 * (dd_ + root), al = 0.5 * (1.0 + ratio);
    const double W[3] = {al * (h * md.bt), al * (h * md.bt), al * (h * md.bn + h * h * md.kn)};
    const double phi[3] = {-(al * md.bt) * pd[0], -(al * md.bt) * pd[1], md.kn * sp - al * (md.bn + h * md.kn) * pd[2]};
    double n[3];
    cross_m(rho, phi, n);
    for (int i = 0; i < 3; ++i) { o[i] -= n[i]; o[3 + i] -= phi[i]; }
    contact_inertia_add(rho, W, Kf);
  }
}
H1_DEV double seq_tau(const DynModel& md, const double* __restrict__ u, int j) {
  if (u == nullptr) return 0.0;
  double tau = u[j - 6];
  const double lo = md.ctrl_lo[j - 6], hi = md.ctrl_hi[j - 6];
  if (tau < lo) tau = lo;
  if (tau > hi) tau = hi;
  return tau;
}

// Per-thread factor / right-hand-side store (local memory; written once and read once per evaluation)
struct SeqFactor {
  double Lm[NV][MAXSLOT];
  double D[NV], rhs[NV];
};

constexpr int seq_roff(int i, int na) { return i * na + i * (i + 1) / 2; }   // offset of chain row i (na + i + 1 entries)

// One serial chain of LEN hinges starting at body b0 whose ancestors are the 6 base dofs and NH hinge dofs
// (motion subspaces Sh): walk down, contact (FOOT), walk up producing the chain's rows of Mhat, then eliminate
// the chain (the steps k = last..first of the L'DL factorisation with fused forward substitution) with every
// index static: the rows, the ancestor block A (packed lower triangle of the (6+NH) x (6+NH) ancestor rows)
// and the right-hand sides stay in registers. Writes L / D / rhs of the chain's dofs to fac.
template <int LEN, int NH, bool FOOT>
H1_DEV void seq_chain(const DynModel& md, int b0, int foot, const double* __restrict__ x, const double* __restrict__ u,
                      const SeqBodyState& parent, const double* Rb, const double* Sh, double* Tpar, double* Kbase,
                      double* A, double* rhsA, SeqFactor& fac) {
  constexpr int NA = 6 + NH;
  const double h = md.h;
  SeqBodyState c = parent;
  double S[LEN][6], own[LEN][16];
  // (NOT unrolled: one copy of the joint / body code per chain instead of LEN — k_primal_factor_seq spills 1.4 instead of 2.7 KB per
  //  thread and has 97 instead of 150 KB of code: 0.99 -> 0.96 ms per 8192-instance launch)
#pragma unroll 1
  for (int i = 0; i < LEN; ++i) {
    seq_joint(md, b0 + i, x, c, S[i]);
    seq_body(md, b0 + i, c, own[i]);
  }
  double Kf[21];
  if (FOOT) {
    seq_contact(md, foot, c, x[2], own[LEN - 1], Kf);
    for (int i = 0; i < 21; ++i) Kbase[i] += Kf[i];
  }
  double rows[LEN * NA + LEN * (LEN + 1) / 2], rc[LEN], T[16];
#pragma unroll
  for (int i = LEN - 1; i >= 0; --i) {
    if (i == LEN - 1) { for (int q = 0; q < 16; ++q) T[q] = own[i][q]; }
    else { for (int q = 0; q < 16; ++q) T[q] += own[i][q]; }
    const int j = 5 + b0 + i;
    const double* Sj = S[i];
    const double bias = Sj[0] * T[0] + Sj[1] * T[1] + Sj[2] * T[2] + Sj[3] * T[3] + Sj[4] * T[4] + Sj[5] * T[5];
    double P[6];
    spi_apply_m(T + 6, Sj, P);
    if (FOOT) contact_inertia_apply(Kf, Sj, P);
    double* row = rows + seq_roff(i, NA);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      row[s] = P[3 + s];
      row[3 + s] = Rb[s] * P[0] + Rb[3 + s] * P[1] + Rb[6 + s] * P[2];
    }
#pragma unroll
    for (int q = 0; q < NH; ++q) {
      const double* Ss = Sh + 6 * q;
      row[6 + q] = Ss[0] * P[0] + Ss[1] * P[1] + Ss[2] * P[2] + Ss[3] * P[3] + Ss[4] * P[4] + Ss[5] * P[5];
    }
#pragma unroll
    for (int s = 0; s <= i; ++s) {
      const double* Ss = S[s];
      row[NA + s] = Ss[0] * P[0] + Ss[1] * P[1] + Ss[2] * P[2] + Ss[3] * P[3] + Ss[4] * P[4] + Ss[5] * P[5];
    }
    row[NA + i] += md.armature[j] + h * md.damping[j];
    rc[i] = seq_tau(md, u, j) - bias - md.damping[j] * x[NQ + j];
  }
  for (int q = 0; q < 16; ++q) Tpar[q] += T[q];
#pragma unroll
  for (int k = LEN - 1; k >= 0; --k) {
    constexpr int dummy = 0; (void)dummy;
    const int n = NA + k + 1;
    double* mk = rows + seq_roff(k, NA);
    const int j = 5 + b0 + k;
    const double dk = mk[n - 1];
    const double inv = 1.0 / dk;
    double a[NA + LEN];
#pragma unroll
    for (int s = 0; s < NA + LEN; ++s) if (s < n - 1) a[s] = mk[s] * inv;
#pragma unroll
    for (int sj = 0; sj < NA + LEN; ++sj) {
      if (sj >= n - 1) continue;
      const double hk = mk[sj];
#pragma unroll
      for (int si = 0; si < NA + LEN; ++si) {
        if (si < sj || si >= n - 1) continue;
        if (si < NA) A[si * (si + 1) / 2 + sj] -= a[si] * hk;
        else rows[seq_roff(si - NA, NA) + sj] -= a[si] * hk;
      }
    }
    const double rk = rc[k];
#pragma unroll
    for (int sj = 0; sj < NA + LEN; ++sj) {
      if (sj >= n - 1) continue;
      fac.Lm[j][sj] = a[sj];
      if (sj < NA) rhsA[sj] -= a[sj] * rk;
      else rc[sj - NA] -= a[sj] * rk;
    }
    fac.D[j] = dk;
    fac.rhs[j] = rk;
  }
}

// back substitution of a chain: acc[k0 + i] = rhs / D - sum L acc[ancestors]; aA = accelerations of the NA ancestors
template <int LEN, int NH>
H1_DEV void seq_chain_back(int k0, const double* aA, const SeqFactor& fac, double* acc) {
  constexpr int NA = 6 + NH;
  double ac[LEN];
#pragma unroll
  for (int i = 0; i < LEN; ++i) {
    const int k = k0 + i;
    double a = fac.rhs[k] / fac.D[k];
#pragma unroll
    for (int s = 0; s < NA; ++s) a -= fac.Lm[k][s] * aA[s];
#pragma unroll
    for (int s = 0; s < i; ++s) a -= fac.Lm[k][NA + s] * ac[s];
    ac[i] = a;
    acc[k] = a;
  }
}

// H1 chain structure the sequential walk is specialised for (validated by build_dyn_model -> DynModel::seq_ok):
// base; bodies 1-5 and 6-10 = leg chains (feet 5, 10); body 11 = torso; bodies 12-15 and 16-19 = arm chains.
constexpr int SEQ_LEG_LEN = 5, SEQ_ARM_LEN = 4, SEQ_TORSO = 2 * SEQ_LEG_LEN + 1;
static_assert(SEQ_TORSO + 2 * SEQ_ARM_LEN == NB - 1, "H1 chain structure");

// x_next = f_D(x, u). x: 51 raw state entries, u: 19 controls (nullptr = zero torques), xn: 51 entries or nullptr.
// pf (optional): factor of Mhat + primal acceleration; com (optional): dynamics-model CoM (world).
H1_DEV void dyn_step_seq(const DynModel& md, const double* __restrict__ x, const double* __restrict__ u,
                         double* __restrict__ xn, PrimalFactor* __restrict__ pf, double* __restrict__ com) {
  SeqFactor fac;
  double A[28], rhsA[7];     // rows of the base dofs (21) and of the torso dof (7), packed lower triangle
  double tot0[16], Kbase[21];
  SeqBodyState base;
  const double h = md.h;
  {
    double qn[4];
    quat_normalize(x + 3, qn);
    quat_to_mat(qn, base.R);
  }
  base.r[0] = base.r[1] = base.r[2] = 0.0;
  for (int i = 0; i < 3; ++i) {
    base.V[i] = base.R[3 * i] * x[NQ + 3] + base.R[3 * i + 1] * x[NQ + 4] + base.R[3 * i + 2] * x[NQ + 5];
    base.V[3 + i] = x[NQ + i];
  }
  {
    double vxw[3];
    cross_m(base.V + 3, base.V, vxw);
    for (int i = 0; i < 3; ++i) { base.Ab[i] = 0.0; base.Ab[3 + i] = vxw[i] - md.gravity[i]; }
  }
  seq_body(md, 0, base, tot0);
  for (int i = 0; i < 21; ++i) Kbase[i] = 0.0;
  for (int i = 0; i < 28; ++i) A[i] = 0.0;
  for (int i = 0; i < 7; ++i) rhsA[i] = 0.0;
  const double* Rb = base.R;
  // legs
#pragma unroll 1
  for (int f = 0; f < 2; ++f)
    seq_chain<SEQ_LEG_LEN, 0, true>(md, 1 + SEQ_LEG_LEN * f, f, x, u, base, Rb, nullptr, tot0, Kbase, A, rhsA, fac);
  // torso (down), arms, torso (up)
  {
    SeqBodyState ts = base;
    double St[6], tott[16];
    seq_joint(md, SEQ_TORSO, x, ts, St);
    seq_body(md, SEQ_TORSO, ts, tott);
#pragma unroll 1
    for (int f = 0; f < 2; ++f)
      seq_chain<SEQ_ARM_LEN, 1, false>(md, SEQ_TORSO + 1 + SEQ_ARM_LEN * f, 0, x, u, ts, Rb, St, tott, nullptr, A, rhsA, fac);
    const int j = 5 + SEQ_TORSO;
    const double bias = St[0] * tott[0] + St[1] * tott[1] + St[2] * tott[2] + St[3] * tott[3] + St[4] * tott[4] + St[5] * tott[5];
    double P[6];
    spi_apply_m(tott + 6, St, P);
    double* row = A + 21;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      row[s] += P[3 + s];
      row[3 + s] += Rb[s] * P[0] + Rb[3 + s] * P[1] + Rb[6 + s] * P[2];
    }
    row[6] += St[0] * P[0] + St[1] * P[1] + St[2] * P[2] + St[3] * P[3] + St[4] * P[4] + St[5] * P[5] +
              md.armature[j] + h * md.damping[j];
    const double rk = rhsA[6] + seq_tau(md, u, j) - bias - md.damping[j] * x[NQ + j];
    for (int q = 0; q < 16; ++q) tot0[q] += tott[q];
    // eliminate the torso dof
    const double dk = row[6], inv = 1.0 / dk;
    double a[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) a[s] = row[s] * inv;
#pragma unroll
    for (int sj = 0; sj < 6; ++sj)
#pragma unroll
      for (int si = sj; si < 6; ++si) A[si * (si + 1) / 2 + sj] -= a[si] * row[sj];
#pragma unroll
    for (int sj = 0; sj < 6; ++sj) { fac.Lm[j][sj] = a[sj]; rhsA[sj] -= a[sj] * rk; }
    fac.D[j] = dk;
    fac.rhs[j] = rk;
  }
  if (com) {
    const double inv = 1.0 / tot0[6];
    com[0] = x[0] + tot0[7] * inv; com[1] = x[1] + tot0[8] * inv; com[2] = x[2] + tot0[9] * inv;
  }
  // base dofs: S_s = [0; e_s] (world-frame linear), [R_base e_{s-3}; 0] (body-frame angular)
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double Sj[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (j < 3) Sj[3 + j] = 1.0;
    else { Sj[0] = Rb[j - 3]; Sj[1] = Rb[3 + j - 3]; Sj[2] = Rb[6 + j - 3]; }
    const double bias = Sj[0] * tot0[0] + Sj[1] * tot0[1] + Sj[2] * tot0[2] + Sj[3] * tot0[3] + Sj[4] * tot0[4] + Sj[5] * tot0[5];
    double P[6];
    spi_apply_m(tot0 + 6, Sj, P);
    contact_inertia_apply(Kbase, Sj, P);
    double* row = A + j * (j + 1) / 2;
#pragma unroll
    for (int s = 0; s <= j; ++s)
      row[s] += (s < 3) ? P[3 + s] : Rb[s - 3] * P[0] + Rb[3 + s - 3] * P[1] + Rb[6 + s - 3] * P[2];
    row[j] += md.armature[j] + h * md.damping[j];
    rhsA[j] += -bias - md.damping[j] * x[NQ + j];
  }
#pragma unroll
  for (int k = 5; k >= 1; --k) {
    double* mk = A + k * (k + 1) / 2;
    const double inv = 1.0 / mk[k];
    double a[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) if (s < k) a[s] = mk[s] * inv;
#pragma unroll
    for (int sj = 0; sj < 5; ++sj) {
      if (sj >= k) continue;
#pragma unroll
      for (int si = 0; si < 5; ++si) {
        if (si < sj || si >= k) continue;
        A[si * (si + 1) / 2 + sj] -= a[si] * mk[sj];
      }
    }
#pragma unroll
    for (int sj = 0; sj < 5; ++sj) {
      if (sj >= k) continue;
      mk[sj] = a[sj];
      rhsA[sj] -= a[sj] * rhsA[k];
    }
  }
  // back substitution, root first
  double acc[NV];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double* mk = A + k * (k + 1) / 2;
    double a = rhsA[k] / mk[k];
#pragma unroll
    for (int s = 0; s < k; ++s) a -= mk[s] * acc[s];
    acc[k] = a;
  }
#pragma unroll 1
  for (int f = 0; f < 2; ++f) seq_chain_back<SEQ_LEG_LEN, 0>(6 + SEQ_LEG_LEN * f, acc, fac, acc);
  {
    double aA[7];
    for (int s = 0; s < 6; ++s) aA[s] = acc[s];
    seq_chain_back<1, 0>(5 + SEQ_TORSO, aA, fac, acc);
    aA[6] = acc[5 + SEQ_TORSO];
#pragma unroll 1
    for (int f = 0; f < 2; ++f) seq_chain_back<SEQ_ARM_LEN, 1>(6 + SEQ_TORSO + SEQ_ARM_LEN * f, aA, fac, acc);
  }
  if (pf) {
    for (int k = 0; k < 6; ++k) {
      const double* mk = A + k * (k + 1) / 2;
      for (int s = 0; s < k; ++s) pf->Lm[k][s] = mk[s];
      pf->D[k] = mk[k];
      pf->a[k] = acc[k];
    }
    for (int k = 6; k < NV; ++k) {
      const int n = md.nlist[k];
      for (int s = 0; s < n - 1; ++s) pf->Lm[k][s] = fac.Lm[k][s];
      pf->D[k] = fac.D[k];
      pf->a[k] = acc[k];
    }
  }
  if (xn) {
    double wn[3];
    for (int j = 0; j < NV; ++j) {
      const double vn = x[NQ + j] + h * acc[j];
      xn[NQ + j] = vn;
      if (j < 3) xn[j] = x[j] + h * vn;
      else if (j < 6) wn[j - 3] = vn;
      else xn[j + 1] = x[j + 1] + h * vn;
    }
    double qo[4];
    quat_step(x + 3, wn, h, qo);
    xn[3] = qo[0]; xn[4] = qo[1]; xn[5] = qo[2]; xn[6] = qo[3];
  }
}

// Dynamics-model CoM only (terminal knot of the line-search cost; RobotUtils::computeCoM,
// /root/reference/src/common/robot_utils.cpp:810-833)
H1_DEV void dyn_com_seq(const DynModel& md, const double* __restrict__ x, double* __restrict__ com) {
  struct Pose { double R[9], r[3]; } cur, saved[SEQ_MAXSAVE];
  {
    double qn[4];
    quat_normalize(x + 3, qn);
    quat_to_mat(qn, cur.R);
  }
  cur.r[0] = cur.r[1] = cur.r[2] = 0.0;
  saved[0] = cur;
  double hsum[3], msum = md.mass[0];
  for (int i = 0; i < 3; ++i)
    hsum[i] = md.mass[0] * (cur.R[3 * i] * md.ipos[0][0] + cur.R[3 * i + 1] * md.ipos[0][1] + cur.R[3 * i + 2] * md.ipos[0][2]);
#pragma unroll 1
  for (int b = 1; b < NB; ++b) {
    const int d = md.depth[b];
    if (md.parent[b] != b - 1) cur = saved[d - 1];
    const double* p = md.pos[b];
    for (int i = 0; i < 3; ++i) cur.r[i] += cur.R[3 * i] * p[0] + cur.R[3 * i + 1] * p[1] + cur.R[3 * i + 2] * p[2];
    if (md.has_rfix[b]) {
      const double* Fx = md.rfix[b];
      double Tm[9];
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k)
          Tm[3 * i + k] = cur.R[3 * i] * Fx[k] + cur.R[3 * i + 1] * Fx[3 + k] + cur.R[3 * i + 2] * Fx[6 + k];
      for (int i = 0; i < 9; ++i) cur.R[i] = Tm[i];
    }
    double sn, cs;
    sincos_t(x[6 + b], &sn, &cs);
    rot_right(cur.R, md.axis[b], sn, cs);
    if (md.nchild[b] > 1) saved[d] = cur;
    const double* ip = md.ipos[b];
    const double m = md.mass[b];
    for (int i = 0; i < 3; ++i)
      hsum[i] += m * (cur.r[i] + cur.R[3 * i] * ip[0] + cur.R[3 * i + 1] * ip[1] + cur.R[3 * i + 2] * ip[2]);
    msum += m;
  }
  const double inv = 1.0 / msum;
  for (int i = 0; i < 3; ++i) com[i] = x[i] + hsum[i] * inv;
}

}  // namespace h1
