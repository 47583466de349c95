// Warp-cooperative cost quadratics of one knot: lx, lu, lxx, luu
// (reference: iLQR::computeCostQuadratics + add*CostDerivatives, /root/reference/src/ilqr/ilqr.cpp:133-244,
//  662-800; the CasADi/Pinocchio derivative providers, /root/reference/src/common/derivatives.cpp:525-707;
//  limit penalties, /root/reference/src/common/robot_utils.cpp:682-778).
//
// The reference differentiates Pinocchio FK with CasADi. Here the exact same gradients and full Hessians
// (all second-order kinematic terms kept, raw un-normalised quaternion coordinates) are produced in closed
// form. Everything is expressed in the BASE frame:
//      P(z) = p_b + R(xi) r(theta)            position of a weighted point set (CoM, or an ankle frame)
//      U(z) = R(xi) u(theta, v),  u = v_b + w_b x r + sum_j r_j thdot_j      its velocity (Pinocchio's
//             LOCAL_WORLD_ALIGNED frame velocity / vcom, with v_b, w_b read as body-frame quantities — Q5)
// with r_j = d r/d theta_j = a_j x mu_j, r_jk = a_j x r_k (j ancestor-or-self of k), and the third-order
// terms needed by the velocity costs folded into D_l and Omega_l (see ph_cq_sets / ph_cq_tables).
// R(xi) is Eigen's quaternion polynomial, so dR/dxi is linear and d2R/dxi2 constant.
// Results are in the Pinocchio coordinate order (quaternion x,y,z,w at 3..6) and are accumulated at the SAME
// indices of the MuJoCo-ordered lx / lxx — reference quirk Q3; the upright term reads x~[3..6] in the roles
// (qw,qx,qy,qz) — quirk Q4. Lane <-> body for the kinematics, lane <-> matrix entries for the assembly.
#pragma once
#include "h1_dyn.cuh"

namespace h1 {

constexpr int CQ_SETS = 3;     // 0: CoM, 1: left ankle frame, 2: right ankle frame
constexpr int CQ_ROWS = 20;    // Jacobian rows: per set P(3) then U(3) -> 18, + 2 balance residual rows
constexpr int CQ_MAXOUTER = 28;
constexpr int CQ_LD = NX + 1;   // leading dimension of the Jacobian rows: = 4 (mod 8) doubles, conflict-free DMMA fragment loads
// TWO warps work on one knot (lanes 0..63 share one CostWarp): the kernel is issue-latency bound per warp and its
// occupancy is set by the 25.8 KB of per-knot scratch, so two warps per scratch block double the warps an SM holds
// (measured: 8 -> 12 resident warps); the phases are loops over columns / joint pairs / tiles that split evenly.
constexpr int CQ_LANES = 64;

struct CostWarp {
  double xt[NX];                 // Pinocchio-ordered state
  double sn[NB], cs[NB];
  double ax[NB][3], o[NB][3], c[NB][3], Om[NB][3];   // hinge axis, body origin, body CoM, relative angular velocity
  double rr[CQ_SETS][3], uu[CQ_SETS][3];
  double rj[CQ_SETS][NB][3], uth[CQ_SETS][NB][3], D[CQ_SETS][NB][3];
  double lamP[CQ_SETS][3], lamU[CQ_SETS][3];          // Hessian-contraction multipliers per set
  double R[9], Ra[4][9];
  double rows[CQ_ROWS][CQ_LD];
  double gcoef[CQ_ROWS];
  double QQ[4][4], QJ[4][NB], JJ[NB][NB], QV[4][NV], JV[NB][NV];
  double gq[4];                  // upright gradient
  double hdiag[NX];              // diagonal additions of lxx: Q (or Qf) and the joint-limit penalty curvature
  int outer_a[CQ_MAXOUTER], outer_b[CQ_MAXOUTER];
  double outer_c[CQ_MAXOUTER];
  int n_outer;
  int bal_on;                    // balance term active at this knot
  double bal_sg, bal_k0, bal_k1; // sigma, U_x sigma', U_y sigma'
};

H1_DEV void cq_dR(const double* xi, int a, double* D) {
  const double x = xi[0], y = xi[1], z = xi[2], w = xi[3];
  if (a == 0) { D[0] = 0; D[1] = 2 * y; D[2] = 2 * z; D[3] = 2 * y; D[4] = -4 * x; D[5] = -2 * w; D[6] = 2 * z; D[7] = 2 * w; D[8] = -4 * x; }
  else if (a == 1) { D[0] = -4 * y; D[1] = 2 * x; D[2] = 2 * w; D[3] = 2 * x; D[4] = 0; D[5] = 2 * z; D[6] = -2 * w; D[7] = 2 * z; D[8] = -4 * y; }
  else if (a == 2) { D[0] = -4 * z; D[1] = -2 * w; D[2] = 2 * x; D[3] = 2 * w; D[4] = -4 * z; D[5] = 2 * y; D[6] = 2 * x; D[7] = 2 * y; D[8] = 0; }
  else { D[0] = 0; D[1] = -2 * z; D[2] = 2 * y; D[3] = 2 * z; D[4] = 0; D[5] = -2 * x; D[6] = -2 * y; D[7] = 2 * x; D[8] = 0; }
}
H1_DEV void mv3(const double* M, const double* v, double* o) {
  const double a = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], b = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
               c = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
H1_DEV void mtv3(const double* M, const double* v, double* o) {
  const double a = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], b = M[1] * v[0] + M[4] * v[1] + M[7] * v[2],
               c = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
H1_DEV bool cq_is_anc(const CostModel& cm, int k, int l) {  // k ancestor-or-self of l
  return cm.depth[k] <= cm.depth[l] && cm.anc_body[l][cm.depth[k]] == k;
}

// ---- phase 0: stage the state in Pinocchio order (convertMuJoCoToPinocchio), sin/cos ----
H1_DEV void ph_cq_load(int lane, CostWarp& w, const double* x) {
  for (int i = lane; i < NX; i += CQ_LANES) {
    int src = i;
    if (i >= 3 && i < 6) src = i + 1; else if (i == 6) src = 3;
    w.xt[i] = x[src];
  }
  if (lane >= 1 && lane < NB) {
    double s, c;
    sincos_t(x[6 + lane], &s, &c);
    w.sn[lane] = s; w.cs[lane] = c;
  }
}

// ---- phase 1: base-frame kinematics, lane <-> body ----
H1_DEV void ph_cq_walk(int lane, const CostModel& cm, CostWarp& w) {
  if (lane >= NB) return;
  const int b = lane;
  double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, r[3] = {0, 0, 0}, Om[3] = {0, 0, 0}, a[3] = {0, 0, 0};
  const int dep = cm.depth[b];
#pragma unroll 1
  for (int d = 1; d <= 5; ++d) {
    if (d <= dep) {
      const int an = cm.anc_body[b][d];
      const double* p = cm.pos[an];
      r[0] += E[0] * p[0] + E[1] * p[1] + E[2] * p[2];
      r[1] += E[3] * p[0] + E[4] * p[1] + E[5] * p[2];
      r[2] += E[6] * p[0] + E[7] * p[1] + E[8] * p[2];
      if (cm.has_rfix[an]) {
        const double* F = cm.rfix[an];
        double T[9];
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k) T[3 * i + k] = E[3 * i] * F[k] + E[3 * i + 1] * F[3 + k] + E[3 * i + 2] * F[6 + k];
        for (int i = 0; i < 9; ++i) E[i] = T[i];
      }
      rot_right(E, cm.axis[an], w.sn[an], w.cs[an]);
      col_of(E, cm.axis[an], a);
      const double thd = w.xt[NQ + 5 + an];
      Om[0] += thd * a[0]; Om[1] += thd * a[1]; Om[2] += thd * a[2];
    }
  }
  const double* ip = cm.ipos[b];
  for (int i = 0; i < 3; ++i) {
    w.ax[b][i] = a[i]; w.o[b][i] = r[i]; w.Om[b][i] = Om[i];
    w.c[b][i] = r[i] + E[3 * i] * ip[0] + E[3 * i + 1] * ip[1] + E[3 * i + 2] * ip[2];
  }
}

// ---- phase 2: per point set, r_l = a_l x mu_l (lane <-> joint l); set centroids / base rotation on lanes 32..35 ----
H1_DEV void ph_cq_sets(int lane, const CostModel& cm, CostWarp& w) {
  if (lane >= 1 && lane < NB) {
    const int l = lane;
    double Hs[3] = {0, 0, 0}, Ws = 0.0;
    for (int i = l; i <= cm.chain_end[l]; ++i) {
      const double m = cm.wmass[i];
      Ws += m; Hs[0] += m * w.c[i][0]; Hs[1] += m * w.c[i][1]; Hs[2] += m * w.c[i][2];
    }
    double mu[3] = {Hs[0] - Ws * w.o[l][0], Hs[1] - Ws * w.o[l][1], Hs[2] - Ws * w.o[l][2]};
    cross3(w.ax[l], mu, w.rj[0][l]);
    for (int f = 0; f < H1_NFOOT; ++f) {
      const int fb = cm.foot_body[f];
      if (cq_is_anc(cm, l, fb)) {
        double m2[3] = {w.o[fb][0] - w.o[l][0], w.o[fb][1] - w.o[l][1], w.o[fb][2] - w.o[l][2]};
        cross3(w.ax[l], m2, w.rj[1 + f][l]);
      } else {
        w.rj[1 + f][l][0] = w.rj[1 + f][l][1] = w.rj[1 + f][l][2] = 0.0;
      }
    }
  }
  if (lane == 32) {
    double s[3] = {0, 0, 0};
    for (int i = 0; i < NB; ++i) { const double m = cm.wmass[i]; s[0] += m * w.c[i][0]; s[1] += m * w.c[i][1]; s[2] += m * w.c[i][2]; }
    w.rr[0][0] = s[0]; w.rr[0][1] = s[1]; w.rr[0][2] = s[2];
  }
  if (lane == 33 || lane == 34) {
    const int f = lane - 33, fb = cm.foot_body[f];
    for (int i = 0; i < 3; ++i) w.rr[1 + f][i] = w.o[fb][i];
  }
  if (lane == 35) {
    const double* xi = &w.xt[3];
    const double x = xi[0], y = xi[1], z = xi[2], q = xi[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * q, twy = ty * q, twz = tz * q, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
                 tyz = tz * y, tzz = tz * z;
    w.R[0] = 1.0 - (tyy + tzz); w.R[1] = txy - twz;         w.R[2] = txz + twy;
    w.R[3] = txy + twz;         w.R[4] = 1.0 - (txx + tzz); w.R[5] = tyz - twx;
    w.R[6] = txz - twy;         w.R[7] = tyz + twx;         w.R[8] = 1.0 - (txx + tyy);
    for (int a = 0; a < 4; ++a) cq_dR(xi, a, w.Ra[a]);
  }
}

// ---- phase 3: D_l, u_theta_l per set (lane <-> joint); set velocities u on lanes 32..34 ----
H1_DEV void ph_cq_vel(int lane, const CostModel& cm, CostWarp& w) {
  const double* vb = &w.xt[NQ];
  const double* wb = &w.xt[NQ + 3];
  if (lane >= 1 && lane < NB) {
    const int l = lane;
    for (int s = 0; s < CQ_SETS; ++s) {
      double D[3] = {0, 0, 0};
      for (int j = l + 1; j <= cm.chain_end[l]; ++j) {
        const double thd = w.xt[NQ + 5 + j];
        D[0] += thd * w.rj[s][j][0]; D[1] += thd * w.rj[s][j][1]; D[2] += thd * w.rj[s][j][2];
      }
      w.D[s][l][0] = D[0]; w.D[s][l][1] = D[1]; w.D[s][l][2] = D[2];
      const double wk[3] = {wb[0] + w.Om[l][0], wb[1] + w.Om[l][1], wb[2] + w.Om[l][2]};
      double t1[3], t2[3];
      cross3(wk, w.rj[s][l], t1);
      cross3(w.ax[l], D, t2);
      w.uth[s][l][0] = t1[0] + t2[0]; w.uth[s][l][1] = t1[1] + t2[1]; w.uth[s][l][2] = t1[2] + t2[2];
    }
  }
  if (lane >= 32 && lane < 32 + CQ_SETS) {
    const int s = lane - 32;
    double acc[3];
    cross3(wb, w.rr[s], acc);
    acc[0] += vb[0]; acc[1] += vb[1]; acc[2] += vb[2];
    for (int j = 1; j < NB; ++j) {
      const double thd = w.xt[NQ + 5 + j];
      acc[0] += thd * w.rj[s][j][0]; acc[1] += thd * w.rj[s][j][1]; acc[2] += thd * w.rj[s][j][2];
    }
    w.uu[s][0] = acc[0]; w.uu[s][1] = acc[1]; w.uu[s][2] = acc[2];
  }
}

// ---- phase 4 (ONE lane, the first of the pair's second warp — the first warp builds the Jacobian rows of phase 5 meanwhile,
//      which need nothing from here): which terms are active at this knot, their multipliers, the rank-1 list ----
struct KnotTargets {
  const double* com_ref;      // [3]
  const double* com_vel_ref;  // [3]
  const double* ee_ref;       // [2][3]
  const int* stance;          // [2]
  bool terminal;
};
// The part of the term setup that needs the state only — the upright term and the zero fills — runs on the SECOND warp during the
// kinematic walk, which occupies the first warp alone (measured, tools/cq_prof.py: 4.0 k cycles of the 43 k of a knot).
H1_DEV void ph_cq_pre(int lane, const H1Weights& wt, CostWarp& w) {
  if (lane > 32) {
    for (int e = lane - 33; e < 2 * CQ_SETS * 3 + CQ_ROWS + 1; e += CQ_LANES - 33) {
      if (e < CQ_SETS * 3) w.lamP[e / 3][e % 3] = 0.0;
      else if (e < 2 * CQ_SETS * 3) w.lamU[(e - CQ_SETS * 3) / 3][e % 3] = 0.0;
      else if (e < 2 * CQ_SETS * 3 + CQ_ROWS) w.gcoef[e - 2 * CQ_SETS * 3] = 0.0;
      else w.bal_on = 0;
    }
    return;
  }
  if (lane != 32) return;
  // upright: closed form on x~[3..6] read as (qw,qx,qy,qz)  (derivatives.cpp:646-666)
  for (int a = 0; a < 4; ++a) { w.gq[a] = 0.0; for (int b = 0; b < 4; ++b) w.QQ[a][b] = 0.0; }
  if (wt.w_upright > 0.0) {
    const double* s = &w.xt[3];
    const double z[3] = {2 * (s[1] * s[3] + s[0] * s[2]), 2 * (s[2] * s[3] - s[0] * s[1]), 1 - 2 * (s[1] * s[1] + s[2] * s[2])};
    const double r[3] = {z[0], z[1], z[2] - 1.0};
    const double J[3][4] = {{2 * s[2], 2 * s[3], 2 * s[0], 2 * s[1]}, {-2 * s[1], -2 * s[0], 2 * s[3], 2 * s[2]}, {0, -4 * s[1], -4 * s[2], 0}};
    const double wu = wt.w_upright;
    for (int a = 0; a < 4; ++a) {
      w.gq[a] = wu * (J[0][a] * r[0] + J[1][a] * r[1] + J[2][a] * r[2]);
      for (int b = 0; b < 4; ++b) w.QQ[a][b] = wu * (J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b]);
    }
    w.QQ[0][2] += wu * r[0] * 2; w.QQ[2][0] += wu * r[0] * 2; w.QQ[1][3] += wu * r[0] * 2; w.QQ[3][1] += wu * r[0] * 2;
    w.QQ[0][1] += wu * r[1] * -2; w.QQ[1][0] += wu * r[1] * -2; w.QQ[2][3] += wu * r[1] * 2; w.QQ[3][2] += wu * r[1] * 2;
    w.QQ[1][1] += wu * r[2] * -4; w.QQ[2][2] += wu * r[2] * -4;
  }
}
H1_DEV void ph_cq_terms(int lane, const H1Weights& wt, const KnotTargets& kt, CostWarp& w) {
  if (lane != 32) return;
  int no = 0;
  double P[CQ_SETS][3], U[CQ_SETS][3];
  for (int s = 0; s < CQ_SETS; ++s) {
    mv3(w.R, w.rr[s], P[s]);
    P[s][0] += w.xt[0]; P[s][1] += w.xt[1]; P[s][2] += w.xt[2];
    mv3(w.R, w.uu[s], U[s]);
  }
  auto sq_term = [&](int s, bool vel, const double* target, double wgt) {
    const int r0 = 6 * s + (vel ? 3 : 0);
    for (int c = 0; c < 3; ++c) {
      const double lam = 2.0 * wgt * ((vel ? U[s][c] : P[s][c]) - target[c]);
      (vel ? w.lamU[s][c] : w.lamP[s][c]) += lam;
      w.gcoef[r0 + c] += lam;
      w.outer_a[no] = r0 + c; w.outer_b[no] = r0 + c; w.outer_c[no] = 2.0 * wgt; ++no;
    }
  };
  const double zero3[3] = {0.0, 0.0, 0.0};
  if (wt.w_com > 0.0) sq_term(0, false, kt.com_ref, wt.w_com);
  if (wt.w_com_vel > 0.0 && !kt.terminal) sq_term(0, true, kt.com_vel_ref, wt.w_com_vel);
  if (wt.w_ee_pos > 0.0)
    for (int f = 0; f < H1_NFOOT; ++f)
      if (kt.stance[f] != 1) sq_term(1 + f, false, kt.ee_ref + 3 * f, wt.w_ee_pos);
  if (wt.w_ee_vel > 0.0)
    for (int f = 0; f < H1_NFOOT; ++f)
      if (kt.stance[f] == 1) sq_term(1 + f, true, zero3, wt.w_ee_vel);
  // balance: 0.5 w || com_xy + vcom_xy sqrt(com_z / 9.81) - p_support ||^2  (derivatives.cpp:668-707)
  if (wt.w_balance > 0.0) {
    double ps[2];
    bool have = true;
    const bool ls = kt.stance[0] == 1, rs = kt.stance[1] == 1;
    const double* l = kt.ee_ref; const double* rf = kt.ee_ref + 3;
    if (ls && rs) { ps[0] = 0.5 * (l[0] + rf[0]); ps[1] = 0.5 * (l[1] + rf[1]); }
    else if (ls) { ps[0] = l[0]; ps[1] = l[1]; }
    else if (rs) { ps[0] = rf[0]; ps[1] = rf[1]; }
    else have = false;
    if (have) {
      const double wb = wt.w_balance, g9 = 9.81;
      const double sg = sqrt(P[0][2] / g9), sg1 = 1.0 / (2.0 * g9 * sg), sg2 = -1.0 / (4.0 * g9 * g9 * sg * sg * sg);
      const double rho0 = P[0][0] + sg * U[0][0] - ps[0], rho1 = P[0][1] + sg * U[0][1] - ps[1];
      // residual rows 18,19 = J_Pk + sg J_Uk + U_k sg1 J_Pz are built in ph_cq_rows2
      w.bal_on = 1; w.bal_sg = sg; w.bal_k0 = U[0][0] * sg1; w.bal_k1 = U[0][1] * sg1;
      w.gcoef[18] = wb * rho0; w.gcoef[19] = wb * rho1;
      w.outer_a[no] = 18; w.outer_b[no] = 18; w.outer_c[no] = wb; ++no;
      w.outer_a[no] = 19; w.outer_b[no] = 19; w.outer_c[no] = wb; ++no;
      w.outer_a[no] = 3; w.outer_b[no] = 2; w.outer_c[no] = wb * rho0 * sg1; ++no;
      w.outer_a[no] = 2; w.outer_b[no] = 3; w.outer_c[no] = wb * rho0 * sg1; ++no;
      w.outer_a[no] = 4; w.outer_b[no] = 2; w.outer_c[no] = wb * rho1 * sg1; ++no;
      w.outer_a[no] = 2; w.outer_b[no] = 4; w.outer_c[no] = wb * rho1 * sg1; ++no;
      w.outer_a[no] = 2; w.outer_b[no] = 2; w.outer_c[no] = wb * (rho0 * U[0][0] + rho1 * U[0][1]) * sg2; ++no;
      w.lamP[0][0] += wb * rho0; w.lamP[0][1] += wb * rho1; w.lamP[0][2] += wb * sg1 * (rho0 * U[0][0] + rho1 * U[0][1]);
      w.lamU[0][0] += wb * sg * rho0; w.lamU[0][1] += wb * sg * rho1;
    }
  }
  w.n_outer = no;
}

// ---- phase 5: Jacobian rows (lane <-> state column) and the contraction tables (lane <-> joint pairs) ----
H1_DEV void ph_cq_rows(int lane, const CostModel& cm, CostWarp& w) {
  // rows: J_P(s) = [I | Ra rr | R r_l | 0],  J_U(s) = [0 | Ra u | R uth_l | R, R(e_m x rr), R r_j]
  // (all 51 columns by the FIRST warp of the pair, two per lane: the second warp's lane 0 runs phase 4 at the same time)
  if (lane >= 32) return;
  for (int i = lane; i < NX; i += 32) {
    for (int s = 0; s < CQ_SETS; ++s) {
      double cp[3] = {0, 0, 0}, cu[3] = {0, 0, 0};
      if (i < 3) { cp[0] = (i == 0) ? 1.0 : 0.0; cp[1] = (i == 1) ? 1.0 : 0.0; cp[2] = (i == 2) ? 1.0 : 0.0; }
      else if (i < 7) { mv3(w.Ra[i - 3], w.rr[s], cp); mv3(w.Ra[i - 3], w.uu[s], cu); }
      else if (i < NQ) { mv3(w.R, w.rj[s][i - 6], cp); mv3(w.R, w.uth[s][i - 6], cu); }
      else if (i < NQ + 3) { const int m = i - NQ; cu[0] = w.R[m]; cu[1] = w.R[3 + m]; cu[2] = w.R[6 + m]; }
      else if (i < NQ + 6) {
        const int m = i - NQ - 3;
        const double e[3] = {m == 0 ? 1.0 : 0.0, m == 1 ? 1.0 : 0.0, m == 2 ? 1.0 : 0.0};
        double er[3];
        cross3(e, w.rr[s], er);
        mv3(w.R, er, cu);
      } else { mv3(w.R, w.rj[s][i - NQ - 5], cu); }
      for (int c = 0; c < 3; ++c) { w.rows[6 * s + c][i] = cp[c]; w.rows[6 * s + 3 + c][i] = cu[c]; }
    }
  }
}
// Ra[a]' lamP[s] and Ra[a]' lamU[s] (4 x 3 x 2 vectors), computed ONCE per knot by 24 lanes of phase 5b instead of by every lane of
// phase 6 for itself (measured, tools/cq_prof.py: the (xi, v) block alone made the second warp's phase 6 half as long again as
// the first's). They live in scratch that is dead between the kinematic walk and the gradient phase: sn / cs and hdiag.
H1_DEV double* cq_RaU(CostWarp& w) { return w.sn; }                 // [4][CQ_SETS][3] = 36 of the 40 doubles of sn, cs
H1_DEV double* cq_RaP(CostWarp& w) { return w.hdiag; }              // 36 of the 51 doubles of hdiag
static_assert(4 * CQ_SETS * 3 <= 2 * NB && 4 * CQ_SETS * 3 <= NX, "scratch for the transformed multipliers");
H1_DEV void ph_cq_rows2(int lane, CostWarp& w) {  // balance residual rows (need the CoM rows complete)
  if (lane < 4 * CQ_SETS * 2) {
    const int a = lane / (2 * CQ_SETS), rem = lane - a * 2 * CQ_SETS, s = rem >> 1;
    if (rem & 1) mtv3(w.Ra[a], w.lamU[s], cq_RaU(w) + (a * CQ_SETS + s) * 3);
    else mtv3(w.Ra[a], w.lamP[s], cq_RaP(w) + (a * CQ_SETS + s) * 3);
  }
  for (int i = lane; i < NX; i += CQ_LANES) {
    double a = 0.0, b = 0.0;
    if (w.bal_on) {
      a = w.rows[0][i] + w.bal_sg * w.rows[3][i] + w.bal_k0 * w.rows[2][i];
      b = w.rows[1][i] + w.bal_sg * w.rows[4][i] + w.bal_k1 * w.rows[2][i];
    }
    w.rows[18][i] = a; w.rows[19][i] = b;
  }
}

H1_DEV void ph_cq_tables(int lane, const CostModel& cm, CostWarp& w) {
  const double* wb = &w.xt[NQ + 3];
  // per-set transformed multipliers (recomputed per lane: cheap, avoids another exchange)
  double RtP[CQ_SETS][3], RtU[CQ_SETS][3];
  for (int s = 0; s < CQ_SETS; ++s) { mtv3(w.R, w.lamP[s], RtP[s]); mtv3(w.R, w.lamU[s], RtU[s]); }
  // (theta_k, theta_l), (theta_k, thdot_l): lane <-> ordered pair index
  // (pairs k <= l from the model's list: the ancestor-or-self pairs come first, the others only store zeros)
  for (int idx = lane; idx < CQ_NPAIRS; idx += CQ_LANES) {
    const int k = cm.pair_k[idx], l = cm.pair_l[idx];
    double jj = 0.0, jv = 0.0;
    if (idx < cm.n_anc_pairs) {
      for (int s = 0; s < CQ_SETS; ++s) {
        double rkl[3], t[3], t2[3], acc[3];
        cross3(w.ax[k], w.rj[s][l], rkl);
        const double wk[3] = {wb[0] + w.Om[k][0], wb[1] + w.Om[k][1], wb[2] + w.Om[k][2]};
        cross3(wk, rkl, acc);
        const double dO[3] = {w.Om[l][0] - w.Om[k][0], w.Om[l][1] - w.Om[k][1], w.Om[l][2] - w.Om[k][2]};
        cross3(dO, w.rj[s][l], t); cross3(w.ax[k], t, t2);
        acc[0] += t2[0]; acc[1] += t2[1]; acc[2] += t2[2];
        cross3(w.ax[l], w.D[s][l], t); cross3(w.ax[k], t, t2);
        acc[0] += t2[0]; acc[1] += t2[1]; acc[2] += t2[2];
        jj += dot3(RtP[s], rkl) + dot3(RtU[s], acc);
        jv += dot3(RtU[s], rkl);
      }
    }
    w.JJ[k][l] = jj; w.JJ[l][k] = jj;
    w.JV[k][5 + l] = jv; w.JV[l][5 + k] = jv;
  }
  // (theta_k, omega_m) and (xi_a, theta_l): lane <-> joint
  if (lane >= 1 && lane < NB) {
    const int k = lane;
    for (int m = 0; m < 3; ++m) {
      const double e[3] = {m == 0 ? 1.0 : 0.0, m == 1 ? 1.0 : 0.0, m == 2 ? 1.0 : 0.0};
      double v = 0.0;
      for (int s = 0; s < CQ_SETS; ++s) { double er[3]; cross3(e, w.rj[s][k], er); v += dot3(RtU[s], er); }
      w.JV[k][3 + m] = v;
      w.JV[k][m] = 0.0;
    }
    for (int a = 0; a < 4; ++a) {
      double v = 0.0;
      for (int s = 0; s < CQ_SETS; ++s)
        v += dot3(cq_RaP(w) + (a * CQ_SETS + s) * 3, w.rj[s][k]) + dot3(cq_RaU(w) + (a * CQ_SETS + s) * 3, w.uth[s][k]);
      w.QJ[a][k] = v;
    }
  }
  // (xi_a, v) : lanes 32..56 <-> velocity entry ; (xi_a, xi_b): lanes 16..31
  if (lane >= 32 && lane < 32 + NV) {
    const int m = lane - 32;
    for (int a = 0; a < 4; ++a) {
      double v = 0.0;
      for (int s = 0; s < CQ_SETS; ++s) {
        const double* RaU = cq_RaU(w) + (a * CQ_SETS + s) * 3;
        if (m < 3) v += RaU[m];
        else if (m < 6) {
          const int mm = m - 3;
          const double e[3] = {mm == 0 ? 1.0 : 0.0, mm == 1 ? 1.0 : 0.0, mm == 2 ? 1.0 : 0.0};
          double er[3];
          cross3(e, w.rr[s], er);
          v += dot3(RaU, er);
        } else v += dot3(RaU, w.rj[s][m - 5]);
      }
      w.QV[a][m] = v;
    }
  }
  if (lane >= 16 && lane < 32) {  // 16 (a,b) pairs on lanes 16..31; d2R/dxi_a dxi_b = dR/dxi_a evaluated at e_b
    const int a = (lane - 16) >> 2, b = (lane - 16) & 3;
    double e[4] = {0, 0, 0, 0};
    if (b == 0) e[0] = 1.0; else if (b == 1) e[1] = 1.0; else if (b == 2) e[2] = 1.0; else e[3] = 1.0;
    double Rab[9];
    cq_dR(e, a, Rab);
    double v = 0.0;
    for (int s = 0; s < CQ_SETS; ++s) {
      double t[3];
      mv3(Rab, w.rr[s], t); v += dot3(w.lamP[s], t);
      mv3(Rab, w.uu[s], t); v += dot3(w.lamU[s], t);
    }
    w.QQ[a][b] += v;  // on top of the upright block written in ph_cq_terms
  }
}

// ---- phase 6: assemble and store lx, lu, lxx, luu (column-major), incl. Q/R tracking and limit penalties ----
H1_DEV double cq_block(const CostWarp& w, int i, int j) {  // contraction part of H(i,j), i >= j
  // classes: [0,3) p, [3,7) xi, [7,26) theta, [26,51) velocity entries
  if (j < 3 || i < 3) return 0.0;
  if (i < 7) return w.QQ[i - 3][j - 3];                       // (xi, xi)
  if (i < NQ) return (j < 7) ? w.QJ[j - 3][i - 6] : w.JJ[i - 6][j - 6];  // (theta, xi) / (theta, theta)
  if (j < 7) return w.QV[j - 3][i - NQ];                      // (v, xi)
  if (j < NQ) return w.JV[j - 6][i - NQ];                     // (v, theta)
  return 0.0;                                                 // (v, v)
}
H1_DEV void limit_d(double val, double lo, double hi, double wgt, double* g, double* h) {
  const double margin = 0.1 * (hi - lo), lo_s = lo + margin, hi_s = hi - margin;
  if (val > hi_s) *g += 2.0 * wgt * (val - hi_s);
  if (val < lo_s) *g += -2.0 * wgt * (lo_s - val);
  if (val > hi_s || val < lo_s) *h += 2.0 * wgt;
}
// gradient lx and the diagonal additions of lxx (kept in w.hdiag for the assembly: no global loads inside the tile loop)
template <bool FULLQ>
H1_DEV void ph_cq_grad(int lane, const DynModel& md, const H1Weights& wt, const double* qo, CostWarp& w, const double* x,
                       const double* x_ref, bool terminal, double* lx) {
  const double* Qd = terminal ? wt.Qfdiag : wt.Qdiag;
  const double* Qo = (FULLQ && qo) ? qo + (terminal ? QOFF_QF : 0) : nullptr;
  for (int i = lane; i < NX; i += CQ_LANES) {
    double g = Qd[i] * (x[i] - x_ref[i]);
    double hd = Qd[i];
    if (FULLQ && Qo) for (int j = 0; j < NX; ++j) g += Qo[j * NX + i] * (x[j] - x_ref[j]);   // lx = Q (x - x_ref) with a full Q (ilqr.cpp:145)
    for (int r = 0; r < CQ_ROWS; ++r)
      if (w.gcoef[r] != 0.0) g += w.gcoef[r] * w.rows[r][i];
    if (i >= 3 && i < 7) g += w.gq[i - 3];
    if (i >= 7 && i < NQ) {
      const double lo = md.jnt_lo[i - 7], hi = md.jnt_hi[i - 7];
      if (isfinite(lo) && isfinite(hi) && lo < hi) limit_d(x[i], lo, hi, wt.w_joint_limits, &g, &hd);
    }
    lx[i] = g;
    w.hdiag[i] = hd;
  }
}
template <bool FULLQ>
H1_DEV void ph_cq_store(int lane, const DynModel& md, const H1Weights& wt, const double* qo, const CostWarp& w, const double* x,
                        const double* u, const double* x_ref, const double* u_ref, bool terminal, double* lx,
                        double* lu, double* lxx, double* luu) {
  const int no = w.n_outer;
  auto diag_terms = [&](int i, double h) { return h + w.hdiag[i]; };
#if defined(__CUDACC__)
  // sum_k c_k rows[a_k] (x) rows[b_k] is the product (rows_a diag(c))' rows_b with the term index k as the
  // contraction dimension: 8 x 8 tiles of the lower triangle on the fp64 tensor core (DMMA m8n8k4), <= 7 k-steps.
  {
    constexpr int KS = (CQ_MAXOUTER + 3) / 4;
    const double* Qo = (FULLQ && qo) ? qo + (terminal ? QOFF_QF : 0) : nullptr;
    const int wl = lane & 31, half = lane >> 5;     // lane within its warp; which of the two warps
    const int g = wl >> 2, t4 = wl & 3, nks = (no + 3) >> 2;
    const double* ra[KS]; const double* rb[KS]; double cc[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = 4 * ks + t4;
      const bool valid = k < no;
      ra[ks] = w.rows[valid ? w.outer_a[k] : 0];
      rb[ks] = w.rows[valid ? w.outer_b[k] : 0];
      cc[ks] = valid ? w.outer_c[k] : 0.0;
    }
#pragma unroll 1
    for (int it = 0; it < (NX + 7) / 8; ++it) {
      const int ia = min(8 * it + g, NX - 1);          // rows 51..55 of the last tile are padding (discarded)
      double af[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) af[ks] = (ks < nks) ? cc[ks] * ra[ks][ia] : 0.0;
#pragma unroll 1
      // tiles of a row strip alternate between the two warps; the strips with an odd tile count give their extra tile to the
      // first warp (strips 0, 6) or to the second (2, 4): 14 tiles each (a plain (it + half) & 1 start made it 16 / 12)
      for (int jt = (half + ((it == 0 || it == 6) ? 0 : 1)) & 1; jt <= it; jt += 2) {
        const int jb = min(8 * jt + g, NX - 1);
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          if (ks < nks) dmma884(c0, c1, af[ks], rb[ks][jb]);
        const int i = 8 * it + g;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int j = 8 * jt + 2 * t4 + q;
          if (i < NX && j <= i) {
            double h = (q ? c1 : c0) + cq_block(w, i, j);
            if (i == j) h = diag_terms(i, h);
            else if (FULLQ && Qo) h += Qo[j * NX + i];                     // lxx = Q + ... with a full Q (ilqr.cpp:149)
            lxx[j * NX + i] = h;      // LOWER triangle only (column-major, rows >= column): the backward pass reads nothing else,
                                      // h1ilqr_get_cost_quadratics mirrors it for the caller; the strided mirror stores were half
                                      // of this kernel's store instructions and 10 KB of its 24 KB of output per knot
          }
        }
      }
    }
  }
#else
  for (int j = 0; j < NX; ++j) {      // lower triangle column by column, mirrored on store
    for (int i = j + lane; i < NX; i += CQ_LANES) {
      double h = cq_block(w, i, j);
      for (int k = 0; k < no; ++k) h += w.outer_c[k] * w.rows[w.outer_a[k]][i] * w.rows[w.outer_b[k]][j];
      if (i == j) h = diag_terms(i, h);
      lxx[j * NX + i] = h;
      lxx[i * NX + j] = h;
    }
  }
#endif
  if (!terminal) {   // luu = diag(R + control-limit curvature): zeros first, then the diagonal by its own lanes
    const double* Ro = (FULLQ && qo) ? qo + QOFF_R : nullptr;
    for (int e = lane; e < NU * NU; e += CQ_LANES) {
      const int i = e % NU, j = e / NU;
      if (i != j) luu[e] = (FULLQ && Ro) ? Ro[e] : 0.0;                     // luu = R + ... (ilqr.cpp:150)
    }
    if (lane >= 32 && lane < 32 + NU) {   // (second warp: the first one has the rank-1 setup of the tile phase above)
      const int i = lane - 32;
      double g = wt.Rdiag[i] * (u[i] - u_ref[i]);
      if (FULLQ && Ro) for (int j = 0; j < NU; ++j) g += Ro[j * NU + i] * (u[j] - u_ref[j]);   // lu = R (u - u_ref) (ilqr.cpp:146)
      double h = wt.Rdiag[i];
      limit_d(u[i], md.ctrl_lo[i], md.ctrl_hi[i], wt.w_control_limits, &g, &h);
      lu[i] = g;
      luu[i * NU + i] = h;
    }
  }
}

#if defined(__CUDACC__)
// the two warps of a knot meet at a named barrier (ids 1.. : one per knot slot of the CTA)
#ifdef CQ_PROF   // debug build only (tools/cq_prof.py): per phase and per warp of the pair, cycles of work before / wait at the pair barrier
__device__ unsigned long long cq_prof_sum[2][16][2];
#define H1_CQ_PHASE(call) { call; const long long t1_ = clock64(); asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory"); \
    const long long t2_ = clock64(); if ((threadIdx.x & 31) == 0) { atomicAdd(&cq_prof_sum[(threadIdx.x >> 5) & 1][cq_ph_][0], (unsigned long long)(t1_ - cq_t_)); \
    atomicAdd(&cq_prof_sum[(threadIdx.x >> 5) & 1][cq_ph_][1], (unsigned long long)(t2_ - t1_)); } cq_t_ = t2_; ++cq_ph_; }
#define H1_CQ_LANE const int lane = threadIdx.x & (CQ_LANES - 1); long long cq_t_ = clock64(); int cq_ph_ = 0;
#else
#define H1_CQ_PHASE(call) { call; asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory"); }
#define H1_CQ_LANE const int lane = threadIdx.x & (CQ_LANES - 1);
#endif
#else
#define H1_CQ_PHASE(call) { for (int lane = 0; lane < CQ_LANES; ++lane) { call; } }
#define H1_CQ_LANE
#endif

template <bool FULLQ = true>   // FULLQ = false: compiled without the full-matrix branches (the kernel used for diagonal weights)
H1_DEV void cost_quadratics_warp(const CostModel& cm, const DynModel& md, const H1Weights& wt, CostWarp& w,
                                 const double* x, const double* u, const double* x_ref, const double* u_ref,
                                 const KnotTargets& kt, double* lx, double* lu, double* lxx, double* luu,
                                 const double* qo = nullptr) {   // qo: off-diagonal parts of full Q / R / Qf (DevWeights::qoff) or nullptr
  H1_CQ_LANE
  H1_CQ_PHASE(ph_cq_load(lane, w, x))
  H1_CQ_PHASE((ph_cq_walk(lane, cm, w), ph_cq_pre(lane, wt, w)))
  H1_CQ_PHASE(ph_cq_sets(lane, cm, w))
  H1_CQ_PHASE(ph_cq_vel(lane, cm, w))
  H1_CQ_PHASE((ph_cq_terms(lane, wt, kt, w), ph_cq_rows(lane, cm, w)))   // measured (tools/cq_prof.py): 5.4 k + 2.5 k cycles one after the other
  H1_CQ_PHASE(ph_cq_rows2(lane, w))
  H1_CQ_PHASE(ph_cq_tables(lane, cm, w))
  H1_CQ_PHASE(ph_cq_grad<FULLQ>(lane, md, wt, qo, w, x, x_ref, kt.terminal, lx))
  H1_CQ_PHASE(ph_cq_store<FULLQ>(lane, md, wt, qo, w, x, u, x_ref, u_ref, kt.terminal, lx, lu, lxx, luu))
}

}  // namespace h1
