// Direction-per-thread analytic linearization of f_D (iLQR::computeLinearization,
// /root/reference/src/ilqr/ilqr.cpp:126-131; replaces RobotUtils::linearizeDynamicsFD,
// /root/reference/src/common/robot_utils.cpp:120-160, by exact tangents).
//
// The warp-cooperative tangent pass of h1_dyn.cuh keeps 25 of 32 lanes busy in its widest phase and far fewer
// in the tree recursions. For batched solves the 70 columns of [A_k | B_k] of every knot are independent, so
// here ONE THREAD owns ONE column and walks the whole kinematic tree sequentially in DFS order; all 32 lanes
// of a warp execute the same instruction stream on different directions (no divergence, no shuffles, no
// shared-memory exchange). With the primal acceleration a held fixed (inverse-dynamics form, see h1_dyn.cuh)
//     t = -d g / d(direction),   g = ID(q, v, a) + (armature + h D) a + D v - tau - sum_i J_i^T F_i ,
//     Mhat adot = t   (two sparse triangular solves with the factor kept by the nominal rollout),
// followed by the tangent of the integrator. Everything is expressed about the base origin in world-aligned
// axes, so subtree wrenches are plain sums: with C_k the cumulative wrench over bodies 1..k (DFS order),
// the subtree wrench of body b is C_{end(b)} - C_{b-1}, and no per-body storage is needed.
// Three instantiations keep the work per direction class minimal:
//   q directions (26): kinematics and velocities carry tangents       -> walk<Dual, Dual>
//   v directions (25): kinematics are plain doubles                    -> walk<double, Dual>
//   u directions (19): t is a unit vector (or 0 when the torque is clamped), solve + integrate only.
// The functions contain no CUDA intrinsics and also compile as plain C++ (tests/emul).
#pragma once
#include "h1_dyn.cuh"

namespace h1 {

constexpr int SEQ_MAXSAVE = 3;  // body states are kept for branch bodies (more than one child) of depth < SEQ_MAXSAVE

template <class TA, class TB, class TO> H1_DEV void cross_m(const TA* a, const TB* b, TO* o) {
  TO x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
// spatial inertia (TI) applied to a motion vector (TM) -> momentum (TM)
template <class TI, class TM> H1_DEV void spi_apply_m(const TI* I, const TM* V, TM* P) {
  const TI m = I[0]; const TI* h = I + 1; const TI* J = I + 4;
  TM hv[3], hw[3];
  cross_m(h, V + 3, hv); cross_m(h, V, hw);
  P[0] = J[0] * V[0] + J[3] * V[1] + J[4] * V[2] + hv[0];
  P[1] = J[3] * V[0] + J[1] * V[1] + J[5] * V[2] + hv[1];
  P[2] = J[4] * V[0] + J[5] * V[1] + J[2] * V[2] + hv[2];
  P[3] = m * V[3] - hw[0];
  P[4] = m * V[4] - hw[1];
  P[5] = m * V[5] - hw[2];
}
// spatial inertia of body b about the base origin in world-aligned axes: m, h(3), I(6: xx yy zz xy xz yz)
template <class TK> H1_DEV void body_inertia_seq(const DynModel& md, int b, const TK* R, const TK* r, TK* I) {
  const double* ip = md.ipos[b];
  TK c[3] = {r[0] + R[0] * ip[0] + R[1] * ip[1] + R[2] * ip[2],
             r[1] + R[3] * ip[0] + R[4] * ip[1] + R[5] * ip[2],
             r[2] + R[6] * ip[0] + R[7] * ip[1] + R[8] * ip[2]};
  const double* J = md.inertia[b];
  TK Tm[9];
  for (int i = 0; i < 3; ++i) {
    Tm[3 * i + 0] = R[3 * i] * J[0] + R[3 * i + 1] * J[3] + R[3 * i + 2] * J[4];
    Tm[3 * i + 1] = R[3 * i] * J[3] + R[3 * i + 1] * J[1] + R[3 * i + 2] * J[5];
    Tm[3 * i + 2] = R[3 * i] * J[4] + R[3 * i + 1] * J[5] + R[3 * i + 2] * J[2];
  }
  const double m = md.mass[b];
  const TK cc = dot3(c, c);
  I[0] = TK(m); I[1] = m * c[0]; I[2] = m * c[1]; I[3] = m * c[2];
  I[4] = dot3(Tm, R) + m * (cc - c[0] * c[0]);
  I[5] = dot3(Tm + 3, R + 3) + m * (cc - c[1] * c[1]);
  I[6] = dot3(Tm + 6, R + 6) + m * (cc - c[2] * c[2]);
  I[7] = dot3(Tm, R + 3) - m * (c[0] * c[1]);
  I[8] = dot3(Tm, R + 6) - m * (c[0] * c[2]);
  I[9] = dot3(Tm + 3, R + 6) - m * (c[1] * c[2]);
}
// F = I (Va + Ab) + V x* (I V)
template <class TK, class TV> H1_DEV void body_wrench_seq(const TK* I, const TV* V, const TV* At, TV* F) {
  TV Ia[6], Iv[6];
  spi_apply_m(I, At, Ia);
  spi_apply_m(I, V, Iv);
  TV a1[3], a2[3], a3[3];
  cross_m(V, Iv, a1); cross_m(V + 3, Iv + 3, a2); cross_m(V, Iv + 3, a3);
  F[0] = Ia[0] + a1[0] + a2[0]; F[1] = Ia[1] + a1[1] + a2[1]; F[2] = Ia[2] + a1[2] + a2[2];
  F[3] = Ia[3] + a3[0]; F[4] = Ia[4] + a3[1]; F[5] = Ia[5] + a3[2];
}

// sin / cos of the hinge of body b: from the per-knot table `sc` (sc[2 (b - 1)], sc[2 (b - 1) + 1]; the batched kernels
// compute it once per knot for all directions) or evaluated here when sc == nullptr. `seeded_here`: the angle carries the
// unit tangent of this direction.
H1_DEV void joint_sincos(const double* __restrict__ x, const double* __restrict__ sc, int b, bool, double* sn, double* cs) {
  if (sc) { *sn = sc[2 * (b - 1)]; *cs = sc[2 * (b - 1) + 1]; }
  else sincos_t(x[6 + b], sn, cs);
}
H1_DEV void joint_sincos(const double* __restrict__ x, const double* __restrict__ sc, int b, bool seeded_here, Dual* sn, Dual* cs) {
  double s, c;
  joint_sincos(x, sc, b, false, &s, &c);
  const double d = seeded_here ? 1.0 : 0.0;
  *sn = Dual(s, c * d); *cs = Dual(c, -s * d);
}

template <class TK, class TV> struct SeqState {
  TK R[9], r[3];
  TV V[6], At[6];   // spatial velocity; total acceleration Va + bias (gravity folded in)
};

// Tangent of the inverse-dynamics residual along input direction `seed` (0..25: q entry, 26..50: v entry),
// t[j] = -d g_j, for the state x (raw, 51 entries) and the fixed primal acceleration a (25 entries).
template <class TK, class TV>
H1_DEV void id_tangent_seq(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a, int seed,
                           double* __restrict__ t, const double* __restrict__ sc = nullptr) {
  SeqState<TK, TV> cur, saved[SEQ_MAXSAVE];
  TK Sst[6][6];      // motion subspaces of the bodies on the current root->body path, by depth
  TV spst[6];        // S_b . C_{b-1} of the same bodies
  TK Vab[6];         // J a of the base body (spatial acceleration without bias)
  {
    TK qr[4], qn[4];
    for (int i = 0; i < 4; ++i) qr[i] = seeded<TK>(x[3 + i], seed == 3 + i, 0.0);
    quat_normalize(qr, qn);
    quat_to_mat(qn, cur.R);
  }
  cur.r[0] = cur.r[1] = cur.r[2] = TK(0.0);
  {
    const TV wb[3] = {seeded<TV>(x[NQ + 3], seed == NQ + 3, 0.0), seeded<TV>(x[NQ + 4], seed == NQ + 4, 0.0),
                      seeded<TV>(x[NQ + 5], seed == NQ + 5, 0.0)};
    for (int i = 0; i < 3; ++i) {
      cur.V[i] = cur.R[3 * i] * wb[0] + cur.R[3 * i + 1] * wb[1] + cur.R[3 * i + 2] * wb[2];
      Vab[i] = cur.R[3 * i] * a[3] + cur.R[3 * i + 1] * a[4] + cur.R[3 * i + 2] * a[5];
      cur.V[3 + i] = seeded<TV>(x[NQ + i], seed == NQ + i, 0.0);
      Vab[3 + i] = TK(a[i]);
    }
    TV vxw[3];
    cross_m(cur.V + 3, cur.V, vxw);
    for (int i = 0; i < 3; ++i) {
      cur.At[i] = TV(Vab[i]);
      cur.At[3 + i] = Vab[3 + i] + vxw[i] - md.gravity[i];
    }
  }
  saved[0] = cur;
  TV F0[6];
  {
    TK I[10];
    body_inertia_seq(md, 0, cur.R, cur.r, I);
    body_wrench_seq(I, cur.V, cur.At, F0);
  }
  const TK qz = seeded<TK>(x[2], seed == 2, 0.0);
  TV C[6];
  for (int i = 0; i < 6; ++i) C[i] = TV(0.0);
#pragma unroll 1
  for (int b = 1; b < NB; ++b) {
    const int d = md.depth[b];
    if (md.parent[b] != b - 1) cur = saved[d - 1];
    {
      const double* p = md.pos[b];
      cur.r[0] += cur.R[0] * p[0] + cur.R[1] * p[1] + cur.R[2] * p[2];
      cur.r[1] += cur.R[3] * p[0] + cur.R[4] * p[1] + cur.R[5] * p[2];
      cur.r[2] += cur.R[6] * p[0] + cur.R[7] * p[1] + cur.R[8] * p[2];
      if (md.has_rfix[b]) {
        const double* Fx = md.rfix[b];
        TK Tm[9];
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k)
            Tm[3 * i + k] = cur.R[3 * i] * Fx[k] + cur.R[3 * i + 1] * Fx[3 + k] + cur.R[3 * i + 2] * Fx[6 + k];
        for (int i = 0; i < 9; ++i) cur.R[i] = Tm[i];
      }
    }
    TK S[6];
    {
      TK sn, cs;
      joint_sincos(x, sc, b, seed == 6 + b, &sn, &cs);
      const int ax = md.axis[b];
      rot_right(cur.R, ax, sn, cs);
      col_of(cur.R, ax, S);
      cross_m(cur.r, S, S + 3);
    }
    {
      const TV vj = seeded<TV>(x[NQ + 5 + b], seed == NQ + 5 + b, 0.0);
      const double aj = a[5 + b];
      for (int i = 0; i < 6; ++i) cur.V[i] += S[i] * vj;
      TV c1[3], c2[3], c3[3];
      cross_m(cur.V, S, c1); cross_m(cur.V, S + 3, c2); cross_m(cur.V + 3, S, c3);
      for (int i = 0; i < 3; ++i) {
        cur.At[i] += S[i] * aj + c1[i] * vj;
        cur.At[3 + i] += S[3 + i] * aj + (c2[i] + c3[i]) * vj;
      }
    }
    if (md.nchild[b] > 1) saved[d] = cur;   // build_dyn_model guarantees d < SEQ_MAXSAVE for branch bodies
    for (int i = 0; i < 6; ++i) Sst[d][i] = S[i];
    spst[d] = S[0] * C[0] + S[1] * C[1] + S[2] * C[2] + S[3] * C[3] + S[4] * C[4] + S[5] * C[5];
    TV F[6];
    {
      TK I[10];
      body_inertia_seq(md, b, cur.R, cur.r, I);
      body_wrench_seq(I, cur.V, cur.At, F);
    }
    for (int f = 0; f < H1_NFOOT; ++f) {
      if (b != md.foot_body[f]) continue;
      TK Va[6];
      for (int i = 0; i < 6; ++i) Va[i] = Vab[i];
      for (int dd = 1; dd <= d; ++dd) {
        const double aa = a[5 + md.anc_body[b][dd]];
        for (int i = 0; i < 6; ++i) Va[i] += Sst[dd][i] * aa;
      }
      const double h = md.h;
#pragma unroll 1
      for (int c = 0; c < H1_NCP; ++c) {
        const double* pt = md.foot_pts[f * H1_NCP + c];
        const TK rho[3] = {cur.r[0] + cur.R[0] * pt[0] + cur.R[1] * pt[1] + cur.R[2] * pt[2],
                           cur.r[1] + cur.R[3] * pt[0] + cur.R[4] * pt[1] + cur.R[5] * pt[2],
                           cur.r[2] + cur.R[6] * pt[0] + cur.R[7] * pt[1] + cur.R[8] * pt[2]};
        TV t1[3]; TK t2[3];
        cross_m(cur.V, rho, t1);
        cross_m(Va, rho, t2);
        const TV pd[3] = {cur.V[3] + t1[0], cur.V[4] + t1[1], cur.V[5] + t1[2]};
        const TK pa[3] = {Va[3] + t2[0], Va[4] + t2[1], Va[5] + t2[2]};
        const TK dd_ = -(qz + rho[2]);
        TK root, ratio;
        root_and_ratio(dd_ * dd_ + md.eps * md.eps, dd_, &root, &ratio);
        const TK sp = 0.5 * (dd_ + root), al = 0.5 * (1.0 + ratio);
        TV Fc[3];
        Fc[0] = -(al * md.bt) * (pd[0] + h * pa[0]);
        Fc[1] = -(al * md.bt) * (pd[1] + h * pa[1]);
        Fc[2] = md.kn * sp - al * ((md.bn + h * md.kn) * pd[2] + (h * md.bn + h * h * md.kn) * pa[2]);
        TV n[3];
        cross_m(rho, Fc, n);
        for (int i = 0; i < 3; ++i) { F[i] -= n[i]; F[3 + i] -= Fc[i]; }
      }
    }
    for (int i = 0; i < 6; ++i) C[i] += F[i];
    // dofs whose subtree ends at this body
    for (int dd = d; dd >= 1; --dd) {
      const int bb = md.anc_body[b][dd];
      if (md.chain_end[bb] != b) break;
      const TV g = Sst[dd][0] * C[0] + Sst[dd][1] * C[1] + Sst[dd][2] * C[2] + Sst[dd][3] * C[3] + Sst[dd][4] * C[4] +
                   Sst[dd][5] * C[5] - spst[dd];
      t[5 + bb] = -(tangent_of(g) + ((seed == NQ + 5 + bb) ? md.damping[5 + bb] : 0.0));
    }
  }
  // base dofs: S_i = e_i (linear) for i < 3, [R_base e_{i-3}; 0] for the body-frame angular dofs
  for (int i = 0; i < 6; ++i) C[i] += F0[i];
  for (int i = 0; i < 3; ++i) {
    t[i] = -(tangent_of(C[3 + i]) + ((seed == NQ + i) ? md.damping[i] : 0.0));
    const TK* R = saved[0].R;
    const TV g = R[i] * C[0] + R[3 + i] * C[1] + R[6 + i] * C[2];
    t[3 + i] = -(tangent_of(g) + ((seed == NQ + 3 + i) ? md.damping[3 + i] : 0.0));
  }
}

// Same tangent for a JOINT direction (angle or rate of the hinge of body bj), exploiting its sparsity: only the
// bodies of subtree(bj) move with the direction, so only they are walked with dual numbers (3.1 of 19 bodies on
// average over H1's hinges); the bodies on the path base -> parent(bj) are walked for their kinematics alone and
// every other body is skipped. t_k is non-zero for the dofs of subtree(bj), for the hinge ancestors of bj and for
// the base: the latter two see the direction only through the total wrench tangent of subtree(bj).
// seed = 6 + bj (angle; TK = Dual) or NQ + 5 + bj (rate; TK = double).
template <class TK, class TV>
H1_DEV void id_tangent_sub(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a, int seed,
                           int bj, double* __restrict__ t, const double* __restrict__ sc = nullptr) {
  SeqState<TK, TV> cur, saved[SEQ_MAXSAVE];
  TK Sst[6][6];
  TV spst[6];
  TK Vab[6];
  for (int j = 0; j < NV; ++j) t[j] = 0.0;
  {
    TK qr[4], qn[4];
    for (int i = 0; i < 4; ++i) qr[i] = TK(x[3 + i]);
    quat_normalize(qr, qn);
    quat_to_mat(qn, cur.R);
  }
  cur.r[0] = cur.r[1] = cur.r[2] = TK(0.0);
  {
    for (int i = 0; i < 3; ++i) {
      cur.V[i] = TV(cur.R[3 * i] * x[NQ + 3] + cur.R[3 * i + 1] * x[NQ + 4] + cur.R[3 * i + 2] * x[NQ + 5]);
      Vab[i] = cur.R[3 * i] * a[3] + cur.R[3 * i + 1] * a[4] + cur.R[3 * i + 2] * a[5];
      cur.V[3 + i] = TV(x[NQ + i]);
      Vab[3 + i] = TK(a[i]);
    }
    TV vxw[3];
    cross_m(cur.V + 3, cur.V, vxw);
    for (int i = 0; i < 3; ++i) {
      cur.At[i] = TV(Vab[i]);
      cur.At[3 + i] = Vab[3 + i] + vxw[i] - md.gravity[i];
    }
  }
  saved[0] = cur;
  const TK qz = TK(x[2]);
  const int dj = md.depth[bj], bend = md.chain_end[bj];
  TV C[6];
  for (int i = 0; i < 6; ++i) C[i] = TV(0.0);
  // bodies on the path above bj (kinematics only), then the bodies of subtree(bj) (DFS-contiguous: bj .. bend)
#pragma unroll 1
  for (int step = 1 - dj; step <= bend - bj; ++step) {
    const bool above = step < 0;
    const int b = above ? md.anc_body[bj][dj + step] : bj + step;
    const int d = md.depth[b];
    if (!above && step > 0 && md.parent[b] != b - 1) cur = saved[d - 1];
    {
      const double* p = md.pos[b];
      cur.r[0] += cur.R[0] * p[0] + cur.R[1] * p[1] + cur.R[2] * p[2];
      cur.r[1] += cur.R[3] * p[0] + cur.R[4] * p[1] + cur.R[5] * p[2];
      cur.r[2] += cur.R[6] * p[0] + cur.R[7] * p[1] + cur.R[8] * p[2];
      if (md.has_rfix[b]) {
        const double* Fx = md.rfix[b];
        TK Tm[9];
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k)
            Tm[3 * i + k] = cur.R[3 * i] * Fx[k] + cur.R[3 * i + 1] * Fx[3 + k] + cur.R[3 * i + 2] * Fx[6 + k];
        for (int i = 0; i < 9; ++i) cur.R[i] = Tm[i];
      }
    }
    TK S[6];
    {
      TK sn, cs;
      joint_sincos(x, sc, b, seed == 6 + b, &sn, &cs);
      const int ax = md.axis[b];
      rot_right(cur.R, ax, sn, cs);
      col_of(cur.R, ax, S);
      cross_m(cur.r, S, S + 3);
    }
    {
      const TV vj = seeded<TV>(x[NQ + 5 + b], seed == NQ + 5 + b, 0.0);
      const double aj = a[5 + b];
      for (int i = 0; i < 6; ++i) cur.V[i] += S[i] * vj;
      TV c1[3], c2[3], c3[3];
      cross_m(cur.V, S, c1); cross_m(cur.V, S + 3, c2); cross_m(cur.V + 3, S, c3);
      for (int i = 0; i < 3; ++i) {
        cur.At[i] += S[i] * aj + c1[i] * vj;
        cur.At[3 + i] += S[3 + i] * aj + (c2[i] + c3[i]) * vj;
      }
    }
    for (int i = 0; i < 6; ++i) Sst[d][i] = S[i];
    if (above) continue;
    if (md.nchild[b] > 1) saved[d] = cur;
    spst[d] = S[0] * C[0] + S[1] * C[1] + S[2] * C[2] + S[3] * C[3] + S[4] * C[4] + S[5] * C[5];
    TV F[6];
    {
      TK I[10];
      body_inertia_seq(md, b, cur.R, cur.r, I);
      body_wrench_seq(I, cur.V, cur.At, F);
    }
    for (int f = 0; f < H1_NFOOT; ++f) {
      if (b != md.foot_body[f]) continue;
      TK Va[6];
      for (int i = 0; i < 6; ++i) Va[i] = Vab[i];
      for (int dd = 1; dd <= d; ++dd) {
        const double aa = a[5 + md.anc_body[b][dd]];
        for (int i = 0; i < 6; ++i) Va[i] += Sst[dd][i] * aa;
      }
      const double h = md.h;
#pragma unroll 1
      for (int c = 0; c < H1_NCP; ++c) {
        const double* pt = md.foot_pts[f * H1_NCP + c];
        const TK rho[3] = {cur.r[0] + cur.R[0] * pt[0] + cur.R[1] * pt[1] + cur.R[2] * pt[2],
                           cur.r[1] + cur.R[3] * pt[0] + cur.R[4] * pt[1] + cur.R[5] * pt[2],
                           cur.r[2] + cur.R[6] * pt[0] + cur.R[7] * pt[1] + cur.R[8] * pt[2]};
        TV t1[3]; TK t2[3];
        cross_m(cur.V, rho, t1);
        cross_m(Va, rho, t2);
        const TV pd[3] = {cur.V[3] + t1[0], cur.V[4] + t1[1], cur.V[5] + t1[2]};
        const TK pa[3] = {Va[3] + t2[0], Va[4] + t2[1], Va[5] + t2[2]};
        const TK dd_ = -(qz + rho[2]);
        TK root, ratio;
        root_and_ratio(dd_ * dd_ + md.eps * md.eps, dd_, &root, &ratio);
        const TK sp = 0.5 * (dd_ + root), al = 0.5 * (1.0 + ratio);
        TV Fc[3];
        Fc[0] = -(al * md.bt) * (pd[0] + h * pa[0]);
        Fc[1] = -(al * md.bt) * (pd[1] + h * pa[1]);
        Fc[2] = md.kn * sp - al * ((md.bn + h * md.kn) * pd[2] + (h * md.bn + h * h * md.kn) * pa[2]);
        TV n[3];
        cross_m(rho, Fc, n);
        for (int i = 0; i < 3; ++i) { F[i] -= n[i]; F[3 + i] -= Fc[i]; }
      }
    }
    for (int i = 0; i < 6; ++i) C[i] += F[i];
    for (int dd = d; dd >= dj; --dd) {   // dofs of subtree(bj) whose subtree ends at this body
      const int bb = md.anc_body[b][dd];
      if (md.chain_end[bb] != b) break;
      const TV g = Sst[dd][0] * C[0] + Sst[dd][1] * C[1] + Sst[dd][2] * C[2] + Sst[dd][3] * C[3] + Sst[dd][4] * C[4] +
                   Sst[dd][5] * C[5] - spst[dd];
      t[5 + bb] = -(tangent_of(g) + ((seed == NQ + 5 + bb) ? md.damping[5 + bb] : 0.0));
    }
  }
  // hinge ancestors of bj and the base see the direction through the wrench tangent of subtree(bj) only
  for (int dd = 1; dd < dj; ++dd) {
    const TV g = Sst[dd][0] * C[0] + Sst[dd][1] * C[1] + Sst[dd][2] * C[2] + Sst[dd][3] * C[3] + Sst[dd][4] * C[4] +
                 Sst[dd][5] * C[5];
    t[5 + md.anc_body[bj][dd]] = -tangent_of(g);
  }
  for (int i = 0; i < 3; ++i) {
    t[i] = -tangent_of(C[3 + i]);
    const TK* R = saved[0].R;
    const TV g = R[i] * C[0] + R[3 + i] * C[1] + R[6 + i] * C[2];
    t[3 + i] = -tangent_of(g);
  }
}

// ------------------------------------------------------------------------------------------------------
// RIGID directions of the base (z, world-frame linear velocity, orientation): the rigid-body part of g does not
// have to be differentiated body by body.
//   * g_rb is invariant under translations and under Galilean boosts (the generalized linear velocity of the base
//     is a world-frame quantity), so along z and along the base linear velocity only the sole contacts change:
//     t = -d(contact part), a walk of the two foot chains with the contact law alone on dual numbers.
//   * a world-frame rotation delta about e_k of the base (body-frame omega / alpha, joint states and the world-
//     frame v_lin / a_lin held fixed) rotates every body pose, motion subspace and body wrench rigidly, except that
//     the uniform field (a_lin - gravity) stays behind:  F_b(delta) = R_delta [F_b + I_b (0; (R_delta' - I) f)],
//     f = a_lin - gravity. With f' = f x e_k and the subtree first moments H_j / masses m_j this gives
//        joints:     d g_j  = S_j . [H_j x f'; m_j f']                      (S_j rotates with its wrench)
//        base omega: d g_3+i = (R e_i) . (H_tot x f')
//        base lin:   d g_i  = (e_k x C_rb,lin + m_tot f')_i                 (its subspace e_i does NOT rotate),
//     C_rb,lin = -(armature + h D) a_lin - D v_lin - C_contact,lin from the residual g = 0 of the primal solve.
//     One plain kinematic walk replaces a dual walk with inertias and wrenches; the contact part is the same
//     foot-chain walk as above with the rotation seeded in R. The four raw-quaternion columns are the linear
//     combinations d theta_k / d quat_i of the three rotation tangents (quat_rot_map).
// ------------------------------------------------------------------------------------------------------

// Contact part of t -= d(S_j . C_contact) along a direction carried by the kinematics of the foot chains (TK),
// the base height qz and the spatial velocity V0 of the base. R0: base rotation, Vab0: J a of the base.
// Clin (optional) accumulates the primal linear contact force sum_feet C_contact,lin.
template <class TK>
H1_DEV void contact_tangent_feet(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a,
                                 const TK* R0, const Dual& qz, const Dual* V0, const TK* Vab0,
                                 double* __restrict__ t, double* __restrict__ Clin, const double* __restrict__ sc = nullptr) {
  TK Sst[6][6];
  const double h = md.h;
#pragma unroll 1
  for (int f = 0; f < H1_NFOOT; ++f) {
    const int fb = md.foot_body[f], df = md.depth[fb];
    TK R[9], r[3], Va[6];
    Dual V[6];
    for (int i = 0; i < 9; ++i) R[i] = R0[i];
    r[0] = r[1] = r[2] = TK(0.0);
    for (int i = 0; i < 6; ++i) { V[i] = V0[i]; Va[i] = Vab0[i]; }
#pragma unroll 1
    for (int dd = 1; dd <= df; ++dd) {
      const int b = md.anc_body[fb][dd];
      const double* p = md.pos[b];
      r[0] += R[0] * p[0] + R[1] * p[1] + R[2] * p[2];
      r[1] += R[3] * p[0] + R[4] * p[1] + R[5] * p[2];
      r[2] += R[6] * p[0] + R[7] * p[1] + R[8] * p[2];
      if (md.has_rfix[b]) {
        const double* Fx = md.rfix[b];
        TK Tm[9];
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k)
            Tm[3 * i + k] = R[3 * i] * Fx[k] + R[3 * i + 1] * Fx[3 + k] + R[3 * i + 2] * Fx[6 + k];
        for (int i = 0; i < 9; ++i) R[i] = Tm[i];
      }
      TK sn, cs, S[6];
      joint_sincos(x, sc, b, false, &sn, &cs);
      const int ax = md.axis[b];
      rot_right(R, ax, sn, cs);
      col_of(R, ax, S);
      cross_m(r, S, S + 3);
      const double vj = x[NQ + 5 + b], aj = a[5 + b];
      for (int i = 0; i < 6; ++i) { V[i] += S[i] * vj; Va[i] += S[i] * aj; Sst[dd][i] = S[i]; }
    }
    Dual F[6];
    for (int i = 0; i < 6; ++i) F[i] = Dual(0.0);
#pragma unroll 1
    for (int c = 0; c < H1_NCP; ++c) {
      const double* pt = md.foot_pts[f * H1_NCP + c];
      const TK rho[3] = {r[0] + R[0] * pt[0] + R[1] * pt[1] + R[2] * pt[2],
                         r[1] + R[3] * pt[0] + R[4] * pt[1] + R[5] * pt[2],
                         r[2] + R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]};
      Dual t1[3]; TK t2[3];
      cross_m(V, rho, t1);
      cross_m(Va, rho, t2);
      const Dual pd[3] = {V[3] + t1[0], V[4] + t1[1], V[5] + t1[2]};
      const TK pa[3] = {Va[3] + t2[0], Va[4] + t2[1], Va[5] + t2[2]};
      const Dual dd_ = -(qz + rho[2]);
      Dual root, ratio;
      root_and_ratio(dd_ * dd_ + md.eps * md.eps, dd_, &root, &ratio);
      const Dual sp = 0.5 * (dd_ + root), al = 0.5 * (1.0 + ratio);
      Dual Fc[3];
      Fc[0] = -(al * md.bt) * (pd[0] + h * pa[0]);
      Fc[1] = -(al * md.bt) * (pd[1] + h * pa[1]);
      Fc[2] = md.kn * sp - al * ((md.bn + h * md.kn) * pd[2] + (h * md.bn + h * h * md.kn) * pa[2]);
      Dual n[3];
      cross_m(rho, Fc, n);
      for (int i = 0; i < 3; ++i) { F[i] -= n[i]; F[3 + i] -= Fc[i]; }
    }
    for (int dd = 1; dd <= df; ++dd) {
      const Dual g = Sst[dd][0] * F[0] + Sst[dd][1] * F[1] + Sst[dd][2] * F[2] + Sst[dd][3] * F[3] + Sst[dd][4] * F[4] +
                     Sst[dd][5] * F[5];
      t[5 + md.anc_body[fb][dd]] -= tangent_of(g);
    }
    for (int i = 0; i < 3; ++i) {
      t[i] -= tangent_of(F[3 + i]);
      const Dual g = R0[i] * F[0] + R0[3 + i] * F[1] + R0[6 + i] * F[2];
      t[3 + i] -= tangent_of(g);
      if (Clin) Clin[i] += val(F[3 + i]);
    }
  }
}

// base rotation (double), spatial velocity of the base and J a of the base from the raw state
H1_DEV void base_frame_seq(const double* __restrict__ x, const double* __restrict__ a, double* R, double* V, double* Vab) {
  double qn[4];
  quat_normalize(x + 3, qn);
  quat_to_mat(qn, R);
  for (int i = 0; i < 3; ++i) {
    V[i] = R[3 * i] * x[NQ + 3] + R[3 * i + 1] * x[NQ + 4] + R[3 * i + 2] * x[NQ + 5];
    Vab[i] = R[3 * i] * a[3] + R[3 * i + 1] * a[4] + R[3 * i + 2] * a[5];
    V[3 + i] = x[NQ + i];
    Vab[3 + i] = a[i];
  }
}

// t = -dg along z (seed 2) or along the world-frame linear velocity of the base (seed NQ + 0..2): contact only.
H1_DEV void id_tangent_rigid(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a, int seed,
                             double* __restrict__ t, const double* __restrict__ sc = nullptr) {
  double R[9], V[6], Vab[6];
  base_frame_seq(x, a, R, V, Vab);
  Dual Vd[6];
  for (int i = 0; i < 6; ++i) Vd[i] = Dual(V[i], (i >= 3 && seed == NQ + i - 3) ? 1.0 : 0.0);
  for (int j = 0; j < NV; ++j) t[j] = 0.0;
  contact_tangent_feet<double>(md, x, a, R, Dual(x[2], seed == 2 ? 1.0 : 0.0), Vd, Vab, t, nullptr, sc);
  if (seed >= NQ) t[seed - NQ] -= md.damping[seed - NQ];
}

// rigid-body part of the rotation tangent: t = -[S_j . (H_j x f'; m_j f')] over one plain kinematic walk
H1_DEV void rot_tangent_rb(const DynModel& md, const double* __restrict__ x, const double* R0, const double* fp,
                           double* __restrict__ t, const double* __restrict__ sc = nullptr) {
  struct Pose { double R[9], r[3]; } cur, saved[SEQ_MAXSAVE];
  double Sst[6][6], spst[6], C[6];
  for (int i = 0; i < 9; ++i) cur.R[i] = R0[i];
  cur.r[0] = cur.r[1] = cur.r[2] = 0.0;
  saved[0] = cur;
  for (int i = 0; i < 6; ++i) C[i] = 0.0;
#pragma unroll 1
  for (int b = 1; b < NB; ++b) {
    const int d = md.depth[b];
    if (md.parent[b] != b - 1) cur = saved[d - 1];
    const double* p = md.pos[b];
    cur.r[0] += cur.R[0] * p[0] + cur.R[1] * p[1] + cur.R[2] * p[2];
    cur.r[1] += cur.R[3] * p[0] + cur.R[4] * p[1] + cur.R[5] * p[2];
    cur.r[2] += cur.R[6] * p[0] + cur.R[7] * p[1] + cur.R[8] * p[2];
    if (md.has_rfix[b]) {
      const double* Fx = md.rfix[b];
      double Tm[9];
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k)
          Tm[3 * i + k] = cur.R[3 * i] * Fx[k] + cur.R[3 * i + 1] * Fx[3 + k] + cur.R[3 * i + 2] * Fx[6 + k];
      for (int i = 0; i < 9; ++i) cur.R[i] = Tm[i];
    }
    double sn, cs, S[6];
    joint_sincos(x, sc, b, false, &sn, &cs);
    const int ax = md.axis[b];
    rot_right(cur.R, ax, sn, cs);
    col_of(cur.R, ax, S);
    cross_m(cur.r, S, S + 3);
    if (md.nchild[b] > 1) saved[d] = cur;
    for (int i = 0; i < 6; ++i) Sst[d][i] = S[i];
    spst[d] = S[0] * C[0] + S[1] * C[1] + S[2] * C[2] + S[3] * C[3] + S[4] * C[4] + S[5] * C[5];
    const double* ip = md.ipos[b];
    const double m = md.mass[b];
    const double hb[3] = {m * (cur.r[0] + cur.R[0] * ip[0] + cur.R[1] * ip[1] + cur.R[2] * ip[2]),
                          m * (cur.r[1] + cur.R[3] * ip[0] + cur.R[4] * ip[1] + cur.R[5] * ip[2]),
                          m * (cur.r[2] + cur.R[6] * ip[0] + cur.R[7] * ip[1] + cur.R[8] * ip[2])};
    double n[3];
    cross_m(hb, fp, n);
    for (int i = 0; i < 3; ++i) { C[i] += n[i]; C[3 + i] += m * fp[i]; }
    for (int dd = d; dd >= 1; --dd) {
      const int bb = md.anc_body[b][dd];
      if (md.chain_end[bb] != b) break;
      t[5 + bb] = -(Sst[dd][0] * C[0] + Sst[dd][1] * C[1] + Sst[dd][2] * C[2] + Sst[dd][3] * C[3] + Sst[dd][4] * C[4] +
                    Sst[dd][5] * C[5] - spst[dd]);
    }
  }
  {
    const double* ip = md.ipos[0];
    const double m = md.mass[0];
    const double hb[3] = {m * (R0[0] * ip[0] + R0[1] * ip[1] + R0[2] * ip[2]), m * (R0[3] * ip[0] + R0[4] * ip[1] + R0[5] * ip[2]),
                          m * (R0[6] * ip[0] + R0[7] * ip[1] + R0[8] * ip[2])};
    double n[3];
    cross_m(hb, fp, n);
    for (int i = 0; i < 3; ++i) { C[i] += n[i]; C[3 + i] += m * fp[i]; }
  }
  for (int i = 0; i < 3; ++i) {
    t[i] = -C[3 + i];
    t[3 + i] = -(R0[i] * C[0] + R0[3 + i] * C[1] + R0[6 + i] * C[2]);
  }
}

// t = -dg / d(theta_k): world-frame rotation of the base about e_k (k = 0, 1, 2) at fixed generalized v, a
H1_DEV void id_tangent_rot(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a, int k,
                           double* __restrict__ t, const double* __restrict__ sc = nullptr) {
  double R[9], V[6], Vab[6];
  base_frame_seq(x, a, R, V, Vab);
  const double ek[3] = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
  const double f0[3] = {a[0] - md.gravity[0], a[1] - md.gravity[1], a[2] - md.gravity[2]};
  double fp[3];
  cross_m(f0, ek, fp);
  rot_tangent_rb(md, x, R, fp, t, sc);
  Dual Rd[9], Vd[6], Vabd[6];
  for (int c = 0; c < 3; ++c) {
    const double v[3] = {R[c], R[3 + c], R[6 + c]};
    double d[3];
    cross_m(ek, v, d);
    for (int i = 0; i < 3; ++i) Rd[3 * i + c] = Dual(R[3 * i + c], d[i]);
  }
  for (int i = 0; i < 3; ++i) {
    Vd[i] = Rd[3 * i] * x[NQ + 3] + Rd[3 * i + 1] * x[NQ + 4] + Rd[3 * i + 2] * x[NQ + 5];
    Vabd[i] = Rd[3 * i] * a[3] + Rd[3 * i + 1] * a[4] + Rd[3 * i + 2] * a[5];
    Vd[3 + i] = Dual(x[NQ + i]);
    Vabd[3 + i] = Dual(a[i]);
  }
  double Clin[3] = {0.0, 0.0, 0.0};
  contact_tangent_feet<Dual>(md, x, a, Rd, Dual(x[2]), Vd, Vabd, t, Clin, sc);
  double crb[3], cx[3];
  for (int i = 0; i < 3; ++i) crb[i] = -(md.armature[i] + md.h * md.damping[i]) * a[i] - md.damping[i] * x[NQ + i] - Clin[i];
  cross_m(ek, crb, cx);
  for (int i = 0; i < 3; ++i) t[i] -= cx[i];
}

// G[k][i] = d theta_k / d quat_raw[i]: world-frame rotation vector of the base per unit change of a raw quaternion
// entry (normalisation included), read off  dR R' = [d theta]x  of the dual-number rotation matrix.
H1_DEV void quat_rot_map(const double* __restrict__ qraw, double G[3][4]) {
  for (int i = 0; i < 4; ++i) {
    Dual q[4], qn[4], Rd[9];
    for (int j = 0; j < 4; ++j) q[j] = Dual(qraw[j], j == i ? 1.0 : 0.0);
    quat_normalize(q, qn);
    quat_to_mat(qn, Rd);
    auto w = [&](int r, int c) { return Rd[3 * r].d * Rd[3 * c].v + Rd[3 * r + 1].d * Rd[3 * c + 1].v + Rd[3 * r + 2].d * Rd[3 * c + 2].v; };
    G[0][i] = w(2, 1); G[1][i] = w(0, 2); G[2][i] = w(1, 0);
  }
}

// Mhat adot = t with the primal factor Mhat = L^T D L (unit-lower rows Lm[k][slot], branch-sparse); in place.
template <bool DINV = false>
H1_DEV void tangent_solve_seq(const DynModel& md, const double* __restrict__ Lm, const double* __restrict__ D,
                              double* __restrict__ t) {
#pragma unroll 1
  for (int k = NV - 1; k >= 1; --k) {
    const int n = md.nlist[k];
    const double tk = t[k];
    for (int s = 0; s < n - 1; ++s) t[md.alist[k][s]] -= Lm[k * MAXSLOT + s] * tk;
  }
#pragma unroll 1
  for (int k = 0; k < NV; ++k) {
    const int n = md.nlist[k];
    double ad = DINV ? t[k] * D[k] : t[k] / D[k];
    for (int s = 0; s < n - 1; ++s) ad -= Lm[k * MAXSLOT + s] * t[md.alist[k][s]];
    t[k] = ad;
  }
}

// Same solve with H1's dof tree compiled in (DynModel::seq_ok): every index is static, t stays in registers.
// DINV: D holds the reciprocals 1 / D_k -> no division per column.
template <bool DINV = false>
H1_DEV void tangent_solve_h1(const double* __restrict__ Lm, const double* __restrict__ D, double* __restrict__ t) {
#pragma unroll
  for (int k = NV - 1; k >= 1; --k) {
    const double tk = t[k];
#pragma unroll
    for (int s = 0; s < MAXSLOT - 1; ++s)
      if (s < h1_nlist(k) - 1) t[h1_anc(k, s)] -= Lm[k * MAXSLOT + s] * tk;
    asm volatile("" ::: "memory");   // keep the loads of row k inside step k (hoisting all 169 of them spills t)
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double ad = DINV ? t[k] * D[k] : t[k] / D[k];
#pragma unroll
    for (int s = 0; s < MAXSLOT - 1; ++s)
      if (s < h1_nlist(k) - 1) ad -= Lm[k * MAXSLOT + s] * t[h1_anc(k, s)];
    t[k] = ad;
    asm volatile("" ::: "memory");
  }
}

// Column of d x_next / d(input) from adot (tangent of the semi-implicit Euler step + quaternion exponential).
// seed: 0..50 state entry, >= 51 control. col: 51 entries, stride 1.
H1_DEV void integrate_tangent_seq(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a,
                                  int seed, const double* __restrict__ adot, double* __restrict__ col) {
  const double h = md.h;
  for (int j = 0; j < NV; ++j) {
    const double vnd = ((seed == NQ + j) ? 1.0 : 0.0) + h * adot[j];
    col[NQ + j] = vnd;
    if (j < 3) col[j] = ((seed == j) ? 1.0 : 0.0) + h * vnd;
    else if (j >= 6) col[j + 1] = ((seed == j + 1) ? 1.0 : 0.0) + h * vnd;
  }
  Dual q[4], wn[3], qo[4];
  for (int i = 0; i < 4; ++i) q[i] = Dual(x[3 + i], (seed == 3 + i) ? 1.0 : 0.0);
  for (int i = 0; i < 3; ++i)
    wn[i] = Dual(x[NQ + 3 + i] + h * a[3 + i], ((seed == NQ + 3 + i) ? 1.0 : 0.0) + h * adot[3 + i]);
  quat_step(q, wn, h, qo);
  for (int i = 0; i < 4; ++i) col[3 + i] = qo[i].d;
}

// Jacobians of the quaternion update of a knot, computed once and shared by its 70 columns:
// J[d][r] = d quat_next[r] / d (raw quaternion entry d) for d < 4, / d (new body-frame angular velocity d - 4) for d >= 4.
constexpr int QJ_DIRS = 7;
H1_DEV void quat_step_jac_dir(const DynModel& md, const double* __restrict__ x, const double* __restrict__ a, int d,
                              double* __restrict__ J4) {
  const double h = md.h;
  Dual q[4], wn[3], qo[4];
  for (int i = 0; i < 4; ++i) q[i] = Dual(x[3 + i], d == i ? 1.0 : 0.0);
  for (int i = 0; i < 3; ++i) wn[i] = Dual(x[NQ + 3 + i] + h * a[3 + i], d == 4 + i ? 1.0 : 0.0);
  quat_step(q, wn, h, qo);
  for (int r = 0; r < 4; ++r) J4[r] = qo[r].d;
}
// integrate_tangent_seq with those Jacobians (J: [QJ_DIRS][4])
H1_DEV void integrate_tangent_pre(const DynModel& md, int seed, const double* __restrict__ adot,
                                  const double* __restrict__ J, double* __restrict__ col) {
  const double h = md.h;
  for (int j = 0; j < NV; ++j) {
    const double vnd = ((seed == NQ + j) ? 1.0 : 0.0) + h * adot[j];
    col[NQ + j] = vnd;
    if (j < 3) col[j] = ((seed == j) ? 1.0 : 0.0) + h * vnd;
    else if (j >= 6) col[j + 1] = ((seed == j + 1) ? 1.0 : 0.0) + h * vnd;
  }
  const double w0 = col[NQ + 3], w1 = col[NQ + 4], w2 = col[NQ + 5];   // d (new body-frame angular velocity)
  for (int r = 0; r < 4; ++r) {
    double v = J[16 + r] * w0 + J[20 + r] * w1 + J[24 + r] * w2;
    if (seed >= 3 && seed < 7) v += J[4 * (seed - 3) + r];
    col[3 + r] = v;
  }
}

// One exact column of [A | B] (seed: 0..50 state entry, 51..69 control) with the cheapest applicable walk.
// H1TREE: the model has H1's dof tree (DynModel::seq_ok) -> static-index triangular solves.
template <bool H1TREE>
H1_DEV void linearize_column_t(const DynModel& md, const double* __restrict__ x, const double* __restrict__ u,
                               const PrimalFactor& pf, int seed, double* __restrict__ col) {
  if (seed < 2) {   // f_D is translation invariant in x and y: the column is the unit vector
    for (int j = 0; j < NX; ++j) col[j] = (j == seed) ? 1.0 : 0.0;
    return;
  }
  double tv[NV];
  if (seed == 2) id_tangent_rigid(md, x, pf.a, seed, tv);                                // z: contact only
  else if (seed < 7) {                                                                   // raw quaternion entry: rotations combined
    double G[3][4], tr[NV];
    quat_rot_map(x + 3, G);
    for (int j = 0; j < NV; ++j) tv[j] = 0.0;
    for (int k = 0; k < 3; ++k) {
      id_tangent_rot(md, x, pf.a, k, tr);
      for (int j = 0; j < NV; ++j) tv[j] += G[k][seed - 3] * tr[j];
    }
  }
  else if (seed < NQ) id_tangent_sub<Dual, Dual>(md, x, pf.a, seed, seed - 6, tv);       // hinge angle
  else if (seed < NQ + 3) id_tangent_rigid(md, x, pf.a, seed, tv);                       // base linear velocity: contact only
  else if (seed < NQ + 6) id_tangent_seq<double, Dual>(md, x, pf.a, seed, tv);           // base angular velocity
  else if (seed < NX) id_tangent_sub<double, Dual>(md, x, pf.a, seed, seed - NQ - 5, tv);  // hinge rate
  else {
    const int j = seed - NX;
    for (int k = 0; k < NV; ++k) tv[k] = 0.0;
    tv[6 + j] = (u[j] < md.ctrl_lo[j] || u[j] > md.ctrl_hi[j]) ? 0.0 : 1.0;   // clamped torque: no sensitivity
  }
  if (H1TREE) tangent_solve_h1(&pf.Lm[0][0], pf.D, tv);
  else tangent_solve_seq(md, &pf.Lm[0][0], pf.D, tv);
  if (seed & 1) integrate_tangent_seq(md, x, pf.a, seed, tv, col);
  else {   // (both integrators are exercised by tests/emul)
    double J[QJ_DIRS * 4];
    for (int d = 0; d < QJ_DIRS; ++d) quat_step_jac_dir(md, x, pf.a, d, J + 4 * d);
    integrate_tangent_pre(md, seed, tv, J, col);
  }
}
H1_DEV void linearize_column(const DynModel& md, const double* x, const double* u, const PrimalFactor& pf, int seed,
                             double* col) {
  if (md.seq_ok) linearize_column_t<true>(md, x, u, pf, seed, col);
  else linearize_column_t<false>(md, x, u, pf, seed, col);
}

}  // namespace h1
