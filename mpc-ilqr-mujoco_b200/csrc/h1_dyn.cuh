// Warp-cooperative evaluation of the one-step dynamics map f_D — the unit every rollout and
// every finite-difference column is made of (reference: RobotUtils::rolloutOneStep,
// /root/reference/src/common/robot_utils.cpp:106-117).
//
// One warp evaluates one f_D. Lane j (0..24) owns dof j; lanes 5..24 also own body (lane-5). The kinematic
// tree (h1::DynModel) is staged in shared memory; per-warp scratch (DynWarp) lives in shared memory too.
//   ph_load     : stage x,u; clamp torques; sin/cos of the hinge angles
//   ph_walk     : every lane walks root->its body (<= 5 hinges): pose, spatial velocity, bias acceleration,
//                 then its body's spatial inertia and bias wrench. No lane waits on another lane here.
//   ph_sums     : subtree sums of inertias/wrenches (bodies are in DFS order => contiguous ranges);
//                 lanes 0..7 evaluate the 8 sole contact points and their Jacobian columns
//   ph_rows     : lane j builds row j of Mhat = CRBA + armature + h*damping + contact terms in the
//                 branch-sparse layout M[j][slot], and rhs_j
//   ph_ltdl     : sparse L^T D L factorisation (Featherstone), one dof per step, fused forward substitution
//   ph_back     : level-synchronous back substitution
//   ph_integrate: semi-implicit Euler + quaternion exponential map
// The phase functions contain no warp intrinsics: between two phases the kernel has a __syncwarp(). The same
// source compiles as plain C++ for the lane-emulation test build (tests/emul), which runs the lanes of a
// phase one after another.
#pragma once
#include "h1_common.cuh"

namespace h1 {

// First-order dual number: value + one directional derivative. Instantiating the assembly phases with
// T = Dual propagates d/d(direction) through f_D exactly (analytic linearization mode).
struct Dual {
  double v, d;
  Dual() = default;  // trivial: Dual lives in unions and shared memory
  H1_HD Dual(double a) : v(a), d(0.0) {}
  H1_HD Dual(double a, double b) : v(a), d(b) {}
};
H1_HD Dual operator+(const Dual& a, const Dual& b) { return Dual(a.v + b.v, a.d + b.d); }
H1_HD Dual operator-(const Dual& a, const Dual& b) { return Dual(a.v - b.v, a.d - b.d); }
H1_HD Dual operator-(const Dual& a) { return Dual(-a.v, -a.d); }
H1_HD Dual operator*(const Dual& a, const Dual& b) { return Dual(a.v * b.v, a.v * b.d + a.d * b.v); }
// (quotients through ONE reciprocal: an fp64 division is ~20 instructions on the device; values within an ulp of the quotient)
H1_HD Dual operator/(const Dual& a, const Dual& b) { const double r = 1.0 / b.v, q = a.v * r; return Dual(q, (a.d - q * b.d) * r); }
H1_HD Dual operator+(const Dual& a, double b) { return Dual(a.v + b, a.d); }
H1_HD Dual operator+(double b, const Dual& a) { return Dual(a.v + b, a.d); }
H1_HD Dual operator-(const Dual& a, double b) { return Dual(a.v - b, a.d); }
H1_HD Dual operator-(double b, const Dual& a) { return Dual(b - a.v, -a.d); }
H1_HD Dual operator*(const Dual& a, double b) { return Dual(a.v * b, a.d * b); }
H1_HD Dual operator*(double b, const Dual& a) { return Dual(a.v * b, a.d * b); }
H1_HD Dual operator/(const Dual& a, double b) { const double r = 1.0 / b; return Dual(a.v * r, a.d * r); }
H1_HD Dual operator/(double b, const Dual& a) { const double r = 1.0 / a.v, q = b * r; return Dual(q, -q * a.d * r); }
H1_HD Dual& operator+=(Dual& a, const Dual& b) { a.v += b.v; a.d += b.d; return a; }
H1_HD Dual& operator-=(Dual& a, const Dual& b) { a.v -= b.v; a.d -= b.d; return a; }
H1_HD bool operator<(const Dual& a, double b) { return a.v < b; }
H1_HD bool operator>(const Dual& a, double b) { return a.v > b; }
H1_HD double sqrt_t(double a) { return ::sqrt(a); }
H1_HD Dual sqrt_t(const Dual& a) { double r = ::sqrt(a.v); return Dual(r, 0.5 * a.d / r); }
// smoothed contact depth: root = sqrt(s2) and ratio = dd / root from one reciprocal square root on the device
H1_HD void root_and_ratio(double s2, double dd, double* root, double* ratio) {
#if defined(__CUDA_ARCH__)
  const double ri = ::rsqrt(s2);
  *root = s2 * ri; *ratio = dd * ri;
#else
  *root = ::sqrt(s2); *ratio = dd / *root;
#endif
}
H1_HD void root_and_ratio(const Dual& s2, const Dual& dd, Dual* root, Dual* ratio) {
#if defined(__CUDA_ARCH__)
  const double ri = ::rsqrt(s2.v);
#else
  const double ri = 1.0 / ::sqrt(s2.v);
#endif
  const double r = s2.v * ri, hd = 0.5 * s2.d * ri;       // d sqrt(s2) = s2' / (2 sqrt(s2))
  *root = Dual(r, hd);
  const double q = dd.v * ri;                              // d (dd / root) = (dd' - q root') / root
  *ratio = Dual(q, (dd.d - q * hd) * ri);
}
H1_HD double tangent_of(double) { return 0.0; }
H1_HD double tangent_of(const Dual& a) { return a.d; }
H1_HD double val(double a) { return a; }
H1_HD double val(const Dual& a) { return a.v; }
H1_HD void sincos_t(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  ::sincos(a, s, c);
#else
  *s = ::sin(a); *c = ::cos(a);
#endif
}
H1_HD void sincos_t(const Dual& a, Dual* s, Dual* c) {
  double sv, cv;
  sincos_t(a.v, &sv, &cv);
  *s = Dual(sv, cv * a.d); *c = Dual(cv, -sv * a.d);
}
// Input entry of f_D: for T = double the seeded entry is shifted by eps (finite-difference column),
// for T = Dual it carries a unit tangent (analytic column).
template <class T> H1_HD T seeded(double value, bool is_seed, double eps);
template <> H1_HD double seeded<double>(double value, bool is_seed, double eps) { return is_seed ? value + eps : value; }
template <> H1_HD Dual seeded<Dual>(double value, bool is_seed, double) { return Dual(value, is_seed ? 1.0 : 0.0); }

template <class T> struct DynWarpT {
  T q[NQ], v[NV], tau[NV];
  T sn[NB], cs[NB];
  T S[NV][6];
  T body[NB][16];            // per body: wrench f(6) [n;l], spatial inertia m, h(3), I(6: xx yy zz xy xz yz)
  T part[4][16];             // subtree totals of the base's children
  T footR[H1_NFOOT][9], footr[H1_NFOOT][3];
  T M[NV][MAXSLOT];          // row j, slot s <-> dof alist[j][s]; the diagonal is slot nlist[j]-1
  T rhs[NV];
  T cp[NCPT][9];             // per contact point: rho(3), W(3), phi(3)
  union {
    T Jc[NCPT][MAXSLOT][3];       // contact Jacobian columns (dead once the rows are assembled)
    double Lm[NV][MAXSLOT];       // unit-lower factor rows (primal pass only)
  };
  T acc[NV];
  T biasv[NV];            // Coriolis + gravity generalized forces (qfrc_bias analogue)
  double tvec[NV];        // tangent right-hand side / intermediate of the tangent solve
  T com[3];                  // dynamics-model CoM (world), by-product used by the line-search cost
};
using DynWarp = DynWarpT<double>;

template <class T> struct DynLaneT {
  T R[9], r[3], V[6], Ab[6], S[6];
  T T16[16];  // subtree totals for the lane's body: wrench(6) + inertia(10)
};
using DynLane = DynLaneT<double>;

template <class T, class U> H1_DEV void cross3(const T* a, const U* b, T* o) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
template <class T> H1_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> H1_DEV T dot6(const T* a, const T* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// spatial inertia I = (m, h, Io) applied to a motion vector [w;v] -> momentum [n;l]
template <class T> H1_DEV void spi_apply(const T* I, const T* V, T* P) {
  const T m = I[0]; const T* h = I + 1; const T* J = I + 4;
  T hv[3], hw[3];
  cross3(h, V + 3, hv); cross3(h, V, hw);
  P[0] = J[0] * V[0] + J[3] * V[1] + J[4] * V[2] + hv[0];
  P[1] = J[3] * V[0] + J[1] * V[1] + J[5] * V[2] + hv[1];
  P[2] = J[4] * V[0] + J[5] * V[1] + J[2] * V[2] + hv[2];
  P[3] = m * V[3] - hw[0];
  P[4] = m * V[4] - hw[1];
  P[5] = m * V[5] - hw[2];
}
// R <- R * Rot(axis, s, c); static indexing only (R lives in registers)
template <class T> H1_DEV void rot_right(T* R, int axis, const T& s, const T& c) {
#define H1_MIX(p, q)                                                        \
  for (int i = 0; i < 3; ++i) {                                             \
    T cp_ = R[3 * i + p], cq_ = R[3 * i + q];                               \
    R[3 * i + p] = c * cp_ + s * cq_;                                       \
    R[3 * i + q] = c * cq_ - s * cp_;                                       \
  }
  if (axis == 0) { H1_MIX(1, 2) } else if (axis == 1) { H1_MIX(2, 0) } else { H1_MIX(0, 1) }
#undef H1_MIX
}
template <class T> H1_DEV void col_of(const T* R, int c, T* o) {
  if (c == 0) { o[0] = R[0]; o[1] = R[3]; o[2] = R[6]; }
  else if (c == 1) { o[0] = R[1]; o[1] = R[4]; o[2] = R[7]; }
  else { o[0] = R[2]; o[1] = R[5]; o[2] = R[8]; }
}
template <class T> H1_DEV void quat_normalize(const T* q, T* qn) {
  T n = sqrt_t(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-12) { qn[0] = T(1.0); qn[1] = qn[2] = qn[3] = T(0.0); }
  else {   // one reciprocal, four products (within an ulp of the four quotients; a quaternion is normalised three times per f_D
           // evaluation and the four fp64 divisions were the top stall line of the line-search kernel)
    const T inv = T(1.0) / n;
    qn[0] = q[0] * inv; qn[1] = q[1] * inv; qn[2] = q[2] * inv; qn[3] = q[3] * inv;
  }
}
template <class T> H1_DEV void quat_to_mat(const T* q, T* R) {
  T q00 = q[0] * q[0], q11 = q[1] * q[1], q22 = q[2] * q[2], q33 = q[3] * q[3];
  T q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q12 = q[1] * q[2], q13 = q[1] * q[3], q23 = q[2] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[1] = 2.0 * (q12 - q03);     R[2] = 2.0 * (q13 + q02);
  R[3] = 2.0 * (q12 + q03);     R[4] = q00 - q11 + q22 - q33; R[5] = 2.0 * (q23 - q01);
  R[6] = 2.0 * (q13 - q02);     R[7] = 2.0 * (q23 + q01);     R[8] = q00 - q11 - q22 + q33;
}

// ---- phase: stage inputs. `seed` selects one input entry (0..50 state, 51..69 control, -1 none): it is
//      shifted by eps when T = double (finite-difference column) or carries the tangent when T = Dual. ----
template <class T, class W>
H1_DEV void ph_load(int lane, const DynModel& md, W& w, const double* x, const double* u, int seed,
                    double eps) {
  if (lane < NQ) w.q[lane] = seeded<T>(x[lane], seed == lane, eps);
  if (lane < NV) {
    w.v[lane] = seeded<T>(x[NQ + lane], seed == NQ + lane, eps);
    T t = T(0.0);
    if (lane >= 6 && u != nullptr) {
      t = seeded<T>(u[lane - 6], seed == NX + lane - 6, eps);
      const double lo = md.ctrl_lo[lane - 6], hi = md.ctrl_hi[lane - 6];
      if (t < lo) t = T(lo);
      if (t > hi) t = T(hi);
    }
    w.tau[lane] = t;
  }
  if (lane >= 1 && lane < NB) {
    T s, c;
    sincos_t(seeded<T>(x[6 + lane], seed == 6 + lane, eps), &s, &c);
    w.sn[lane] = s; w.cs[lane] = c;
  }
}

// ---- phase: root->body walk, body inertia and bias wrench ----
template <class T> H1_DEV void ph_walk(int lane, const DynModel& md, DynWarpT<T>& w, DynLaneT<T>& L) {
  if (lane >= NV) return;
  const int b = lane < 6 ? 0 : lane - 5;
  T qn[4];
  quat_normalize(&w.q[3], qn);
  quat_to_mat(qn, L.R);
  L.r[0] = L.r[1] = L.r[2] = T(0.0);
  const T wb[3] = {w.v[3], w.v[4], w.v[5]};
  L.V[0] = L.R[0] * wb[0] + L.R[1] * wb[1] + L.R[2] * wb[2];
  L.V[1] = L.R[3] * wb[0] + L.R[4] * wb[1] + L.R[5] * wb[2];
  L.V[2] = L.R[6] * wb[0] + L.R[7] * wb[1] + L.R[8] * wb[2];
  L.V[3] = w.v[0]; L.V[4] = w.v[1]; L.V[5] = w.v[2];
  {
    T vxw[3];
    cross3(L.V + 3, L.V, vxw);  // v_lin x omega_world
    L.Ab[0] = L.Ab[1] = L.Ab[2] = T(0.0);
    L.Ab[3] = vxw[0] - md.gravity[0]; L.Ab[4] = vxw[1] - md.gravity[1]; L.Ab[5] = vxw[2] - md.gravity[2];
  }
  // own motion subspace for the base dofs
  for (int i = 0; i < 6; ++i) L.S[i] = T(0.0);
  if (lane < 3) { L.S[3] = T(lane == 0 ? 1.0 : 0.0); L.S[4] = T(lane == 1 ? 1.0 : 0.0); L.S[5] = T(lane == 2 ? 1.0 : 0.0); }
  else if (lane < 6) { col_of(L.R, lane - 3, L.S); }
  const int dep = md.depth[b];
#pragma unroll 1
  for (int d = 1; d <= 5; ++d) {
    if (d <= dep) {
      const int a = md.anc_body[b][d];
      const double* p = md.pos[a];
      L.r[0] += L.R[0] * p[0] + L.R[1] * p[1] + L.R[2] * p[2];
      L.r[1] += L.R[3] * p[0] + L.R[4] * p[1] + L.R[5] * p[2];
      L.r[2] += L.R[6] * p[0] + L.R[7] * p[1] + L.R[8] * p[2];
      if (md.has_rfix[a]) {
        const double* F = md.rfix[a];
        T Tm[9];
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k)
            Tm[3 * i + k] = L.R[3 * i] * F[k] + L.R[3 * i + 1] * F[3 + k] + L.R[3 * i + 2] * F[6 + k];
        for (int i = 0; i < 9; ++i) L.R[i] = Tm[i];
      }
      const int ax = md.axis[a];
      rot_right(L.R, ax, w.sn[a], w.cs[a]);
      col_of(L.R, ax, L.S);
      cross3(L.r, L.S, L.S + 3);
      const T vj = w.v[5 + a];
      for (int i = 0; i < 6; ++i) L.V[i] += L.S[i] * vj;
      T c1[3], c2[3], c3[3];
      cross3(L.V, L.S, c1); cross3(L.V, L.S + 3, c2); cross3(L.V + 3, L.S, c3);
      L.Ab[0] += c1[0] * vj; L.Ab[1] += c1[1] * vj; L.Ab[2] += c1[2] * vj;
      L.Ab[3] += (c2[0] + c3[0]) * vj; L.Ab[4] += (c2[1] + c3[1]) * vj; L.Ab[5] += (c2[2] + c3[2]) * vj;
    }
  }
  for (int i = 0; i < 6; ++i) w.S[lane][i] = L.S[i];
  if (lane < 5) return;  // lanes 0..4 share body 0 with lane 5
  for (int f = 0; f < H1_NFOOT; ++f)
    if (b == md.foot_body[f]) {
      for (int i = 0; i < 9; ++i) w.footR[f][i] = L.R[i];
      for (int i = 0; i < 3; ++i) w.footr[f][i] = L.r[i];
    }
  // spatial inertia about the base origin, world-aligned
  T I[10];
  {
    const double* ip = md.ipos[b];
    T c[3] = {L.r[0] + L.R[0] * ip[0] + L.R[1] * ip[1] + L.R[2] * ip[2],
              L.r[1] + L.R[3] * ip[0] + L.R[4] * ip[1] + L.R[5] * ip[2],
              L.r[2] + L.R[6] * ip[0] + L.R[7] * ip[1] + L.R[8] * ip[2]};
    const double* J = md.inertia[b];
    T Tm[9];
    for (int i = 0; i < 3; ++i) {
      Tm[3 * i + 0] = L.R[3 * i] * J[0] + L.R[3 * i + 1] * J[3] + L.R[3 * i + 2] * J[4];
      Tm[3 * i + 1] = L.R[3 * i] * J[3] + L.R[3 * i + 1] * J[1] + L.R[3 * i + 2] * J[5];
      Tm[3 * i + 2] = L.R[3 * i] * J[4] + L.R[3 * i + 1] * J[5] + L.R[3 * i + 2] * J[2];
    }
    const double m = md.mass[b];
    const T cc = dot3(c, c);
    I[0] = T(m); I[1] = m * c[0]; I[2] = m * c[1]; I[3] = m * c[2];
    I[4] = dot3(Tm, L.R) + m * (cc - c[0] * c[0]);
    I[5] = dot3(Tm + 3, L.R + 3) + m * (cc - c[1] * c[1]);
    I[6] = dot3(Tm + 6, L.R + 6) + m * (cc - c[2] * c[2]);
    I[7] = dot3(Tm, L.R + 3) - m * (c[0] * c[1]);
    I[8] = dot3(Tm, L.R + 6) - m * (c[0] * c[2]);
    I[9] = dot3(Tm + 3, L.R + 6) - m * (c[1] * c[2]);
  }
  T Ia[6], Iv[6];
  spi_apply(I, L.Ab, Ia);
  spi_apply(I, L.V, Iv);
  {
    T a1[3], a2[3], a3[3];
    cross3(L.V, Iv, a1); cross3(L.V + 3, Iv + 3, a2); cross3(L.V, Iv + 3, a3);
    w.body[b][0] = Ia[0] + a1[0] + a2[0]; w.body[b][1] = Ia[1] + a1[1] + a2[1]; w.body[b][2] = Ia[2] + a1[2] + a2[2];
    w.body[b][3] = Ia[3] + a3[0]; w.body[b][4] = Ia[4] + a3[1]; w.body[b][5] = Ia[5] + a3[2];
  }
  for (int i = 0; i < 10; ++i) w.body[b][6 + i] = I[i];
}

// ---- phase: subtree sums for non-base bodies; contact points on lanes 0..7 ----
template <class T> H1_DEV void ph_sums(int lane, const DynModel& md, DynWarpT<T>& w, DynLaneT<T>& L) {
  if (lane >= 6 && lane < NV) {
    const int b = lane - 5, e = md.chain_end[b];
    for (int i = 0; i < 16; ++i) L.T16[i] = w.body[b][i];
#pragma unroll 1
    for (int k = b + 1; k <= e; ++k)
      for (int i = 0; i < 16; ++i) L.T16[i] += w.body[k][i];
    if (md.parent[b] == 0) {
      const int slot = md.base_child_slot[b];
      for (int i = 0; i < 16; ++i) w.part[slot][i] = L.T16[i];
    }
  }
  if (lane < NCPT) {
    const int f = lane / H1_NCP;
    const T* R = w.footR[f];
    const double* pt = md.foot_pts[lane];
    T rho[3] = {w.footr[f][0] + R[0] * pt[0] + R[1] * pt[1] + R[2] * pt[2],
                w.footr[f][1] + R[3] * pt[0] + R[4] * pt[1] + R[5] * pt[2],
                w.footr[f][2] + R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]};
    T pd[3] = {T(0.0), T(0.0), T(0.0)};
    const int fd = md.foot_dof[f];
#pragma unroll 1
    for (int s = 0; s < MAXSLOT; ++s) {
      const int k = md.alist[fd][s];
      const T* Sk = w.S[k];
      T t[3];
      cross3(Sk, rho, t);
      const T c0 = Sk[3] + t[0], c1 = Sk[4] + t[1], c2 = Sk[5] + t[2];
      w.Jc[lane][s][0] = c0; w.Jc[lane][s][1] = c1; w.Jc[lane][s][2] = c2;
      const T vk = w.v[k];
      pd[0] += c0 * vk; pd[1] += c1 * vk; pd[2] += c2 * vk;
    }
    const double h = md.h;
    const T dd = -(w.q[2] + rho[2]);
    const T root = sqrt_t(dd * dd + md.eps * md.eps);
    const T sp = 0.5 * (dd + root), al = 0.5 * (1.0 + dd / root);
    w.cp[lane][0] = rho[0]; w.cp[lane][1] = rho[1]; w.cp[lane][2] = rho[2];
    w.cp[lane][3] = al * (h * md.bt); w.cp[lane][4] = al * (h * md.bt); w.cp[lane][5] = al * (h * md.bn + h * h * md.kn);
    w.cp[lane][6] = -(al * md.bt) * pd[0];
    w.cp[lane][7] = -(al * md.bt) * pd[1];
    w.cp[lane][8] = md.kn * sp - al * (md.bn + h * md.kn) * pd[2];
  }
}

// ---- phase: row j of Mhat and rhs_j ----
template <class T> H1_DEV void ph_rows(int lane, const DynModel& md, DynWarpT<T>& w, DynLaneT<T>& L) {
  if (lane >= NV) return;
  if (lane < 6) {
    for (int i = 0; i < 16; ++i) L.T16[i] = w.body[0][i];
    for (int c = 0; c < md.n_base_children; ++c)
      for (int i = 0; i < 16; ++i) L.T16[i] += w.part[c][i];
    if (lane == 0) {
      const T inv = 1.0 / L.T16[6];
      w.com[0] = w.q[0] + L.T16[7] * inv; w.com[1] = w.q[1] + L.T16[8] * inv; w.com[2] = w.q[2] + L.T16[9] * inv;
    }
  }
  const int j = lane, n = md.nlist[j];
  const T bias = dot6(L.S, L.T16);
  T P[6];
  spi_apply(L.T16 + 6, L.S, P);
  T row[MAXSLOT];
#pragma unroll
  for (int s = 0; s < MAXSLOT; ++s) row[s] = (s < n) ? dot6(w.S[md.alist[j][s]], P) : T(0.0);
  w.biasv[j] = bias;
  T rhs = w.tau[j] - bias - md.damping[j] * w.v[j];
  const double diag_add = md.armature[j] + md.h * md.damping[j];
  // contact points that move with dof j
  const int self = n - 1;
#pragma unroll 1
  for (int p = md.cp_lo[j]; p < md.cp_hi[j]; ++p) {
    const T* W = &w.cp[p][3];
    const T* ph = &w.cp[p][6];
    const T* Jj = w.Jc[p][self];
    const T wj0 = W[0] * Jj[0], wj1 = W[1] * Jj[1], wj2 = W[2] * Jj[2];
    rhs += Jj[0] * ph[0] + Jj[1] * ph[1] + Jj[2] * ph[2];
#pragma unroll
    for (int s = 0; s < MAXSLOT; ++s)
      if (s < n) row[s] += wj0 * w.Jc[p][s][0] + wj1 * w.Jc[p][s][1] + wj2 * w.Jc[p][s][2];
  }
#pragma unroll
  for (int s = 0; s < MAXSLOT; ++s)
    if (s < n) w.M[j][s] = (s == self) ? row[s] + diag_add : row[s];
  w.rhs[j] = rhs;
}

// ---- phase: one step (dof k) of the sparse L^T D L factorisation with fused forward substitution ----
// lane = slot sj of row k. Rows of ancestors are updated; row k is normalised into Lm[k] (Lm aliases Jc,
// which ph_rows has finished reading before the first ph_ltdl call).
H1_DEV void ph_ltdl(int lane, int k, const DynModel& md, DynWarp& w) {
  const int n = md.nlist[k];
  const int sj = lane;
  if (sj >= n - 1) return;
  const double inv = 1.0 / w.M[k][n - 1];
  const double hk = w.M[k][sj];
  for (int si = sj; si < n - 1; ++si) {
    const double a = w.M[k][si] * inv;
    w.M[md.alist[k][si]][sj] -= a * hk;
  }
  const double lk = hk * inv;
  w.Lm[k][sj] = lk;
  w.rhs[md.alist[k][sj]] -= lk * w.rhs[k];
}

// ---- phase: back substitution, one level of the dof tree per call (root first) ----
H1_DEV void ph_back(int lane, int lev, const DynModel& md, DynWarp& w) {
  if (lane >= NV || md.level[lane] != lev) return;
  const int k = lane, n = md.nlist[k];
  double a = w.rhs[k] / w.M[k][n - 1];
  for (int s = 0; s < n - 1; ++s) a -= w.Lm[k][s] * w.acc[md.alist[k][s]];
  w.acc[k] = a;
}

// q_next = normalize(normalize(q) * exp(h * wn / 2)) for the body-frame angular velocity wn (quaternion w,x,y,z)
template <class T> H1_DEV void quat_step(const T* q, const T* wn, double h, T* qo) {
  T qu[4];
  quat_normalize(q, qu);
  const T ph[3] = {h * wn[0], h * wn[1], h * wn[2]};
  const T ang = sqrt_t(ph[0] * ph[0] + ph[1] * ph[1] + ph[2] * ph[2]);
  T e[4];
  if (ang < 1e-10) { e[0] = T(1.0); e[1] = 0.5 * ph[0]; e[2] = 0.5 * ph[1]; e[3] = 0.5 * ph[2]; }
  else {
    T s, c;
    sincos_t(0.5 * ang, &s, &c);
    const T sc = s / ang;
    e[0] = c; e[1] = sc * ph[0]; e[2] = sc * ph[1]; e[3] = sc * ph[2];
  }
  const T pq[4] = {qu[0] * e[0] - qu[1] * e[1] - qu[2] * e[2] - qu[3] * e[3],
                   qu[0] * e[1] + qu[1] * e[0] + qu[2] * e[3] - qu[3] * e[2],
                   qu[0] * e[2] - qu[1] * e[3] + qu[2] * e[0] + qu[3] * e[1],
                   qu[0] * e[3] + qu[1] * e[2] - qu[2] * e[1] + qu[3] * e[0]};
  quat_normalize(pq, qo);
}

// ---- phase: integrate and store x_next (T = double) or its tangent (T = Dual, stores .d) ----
H1_HD void store_out(double* p, double a) { *p = a; }
H1_HD void store_out(double* p, const Dual& a) { *p = a.d; }
template <class T, class W> H1_DEV void ph_integrate(int lane, const DynModel& md, W& w, double* xn) {
  if (lane >= NV) return;
  const double h = md.h;
  const T vn = w.v[lane] + h * w.acc[lane];
  store_out(&xn[NQ + lane], vn);
  if (lane < 3) store_out(&xn[lane], w.q[lane] + h * vn);
  else if (lane >= 6) store_out(&xn[lane + 1], w.q[lane + 1] + h * vn);
  else if (lane == 3) {
    const T wn[3] = {w.v[3] + h * w.acc[3], w.v[4] + h * w.acc[4], w.v[5] + h * w.acc[5]};
    T qo[4];
    quat_step(&w.q[3], wn, h, qo);
    store_out(&xn[3], qo[0]); store_out(&xn[4], qo[1]); store_out(&xn[5], qo[2]); store_out(&xn[6], qo[3]);
  }
}

// The sequence of phases. H1_PHASE(call) runs `call` for every lane and then synchronises the warp.
#if defined(__CUDACC__)
#define H1_PHASE(call) { call; __syncwarp(); }
#define H1_LANES_DECL(T) DynLaneT<T> Lr; const int lane = threadIdx.x & 31;
#define H1_LR Lr
#else
#define H1_PHASE(call) { for (int lane = 0; lane < 32; ++lane) { call; } }
#define H1_LANES_DECL(T) static thread_local DynLaneT<T> Lr_[32];
#define H1_LR Lr_[lane]
#endif

// x_next = f_D(x, u). x, u, xn: global or shared memory.
H1_DEV void dyn_step_warp(const DynModel& md, DynWarp& w, const double* x, const double* u, double* xn,
                          int seed = -1, double eps = 0.0) {
  H1_LANES_DECL(double)
  H1_PHASE((ph_load<double, DynWarp>(lane, md, w, x, u, seed, eps)))
  H1_PHASE(ph_walk<double>(lane, md, w, H1_LR))
  H1_PHASE(ph_sums<double>(lane, md, w, H1_LR))
  H1_PHASE(ph_rows<double>(lane, md, w, H1_LR))
  for (int k = NV - 1; k >= 1; --k) H1_PHASE(ph_ltdl(lane, k, md, w))
  for (int lev = 0; lev < MAXSLOT; ++lev) H1_PHASE(ph_back(lane, lev, md, w))
  H1_PHASE((ph_integrate<double, DynWarp>(lane, md, w, xn)))
}

// Assembly only: fills w.com, w.footr/footR, w.M and w.rhs (= tau - bias - damping*v + contact).
H1_DEV void dyn_assemble_warp(const DynModel& md, DynWarp& w, const double* x, const double* u) {
  H1_LANES_DECL(double)
  H1_PHASE((ph_load<double, DynWarp>(lane, md, w, x, u, -1, 0.0)))
  H1_PHASE(ph_walk<double>(lane, md, w, H1_LR))
  H1_PHASE(ph_sums<double>(lane, md, w, H1_LR))
  H1_PHASE(ph_rows<double>(lane, md, w, H1_LR))
}

// ------------------------------------------------------------------------------------------------------
// Analytic linearization: exact columns of A = d f_D / dx and B = d f_D / du by forward-mode tangents.
// The primal pass factorises Mhat once per knot (PrimalFactor); every tangent direction then needs only the
// assembly phases on dual numbers and two sparse triangular solves with the SAME factor:
//      Mhat a = rhs   =>   Mhat adot = rhsdot - Mhatdot a .
// ------------------------------------------------------------------------------------------------------
struct PrimalFactor {
  double Lm[NV][MAXSLOT];
  double D[NV];
  double a[NV];
};

H1_DEV void ph_save_factor(int lane, const DynModel& md, const DynWarp& w, PrimalFactor& pf) {
  if (lane >= NV) return;
  const int n = md.nlist[lane];
  for (int s = 0; s < n - 1; ++s) pf.Lm[lane][s] = w.Lm[lane][s];
  pf.D[lane] = w.M[lane][n - 1];
  pf.a[lane] = w.acc[lane];
}
// t_j = rhsdot_j - (Mhatdot a)_j  with Mhatdot in the branch-sparse row/slot layout (symmetric)
H1_DEV void ph_tan_rhs(int lane, const DynModel& md, DynWarpT<Dual>& w, const PrimalFactor& pf) {
  if (lane >= NV) return;
  const int j = lane, n = md.nlist[j];
  double t = w.rhs[j].d;
  for (int s = 0; s < n; ++s) t -= w.M[j][s].d * pf.a[md.alist[j][s]];
  const int e = md.dof_sub_end[j];
  for (int k = j + 1; k <= e; ++k) t -= w.M[k][n - 1].d * pf.a[k];
  w.tvec[j] = t;
}
// forward substitution L^T y = t, one level per call, leaves first (gather over the dof subtree)
H1_DEV void ph_tan_fwd(int lane, int lev, const DynModel& md, DynWarpT<Dual>& w, const PrimalFactor& pf) {
  if (lane >= NV || md.level[lane] != lev) return;
  const int i = lane, slot = md.nlist[i] - 1, e = md.dof_sub_end[i];
  double y = w.tvec[i];
  for (int k = i + 1; k <= e; ++k) y -= pf.Lm[k][slot] * w.tvec[k];
  w.tvec[i] = y;
}
// back substitution L adot = D^-1 y, one level per call, root first; result into acc = (a, adot)
H1_DEV void ph_tan_back(int lane, int lev, const DynModel& md, DynWarpT<Dual>& w, const PrimalFactor& pf) {
  if (lane >= NV || md.level[lane] != lev) return;
  const int k = lane, n = md.nlist[k];
  double ad = w.tvec[k] / pf.D[k];
  for (int s = 0; s < n - 1; ++s) ad -= pf.Lm[k][s] * w.acc[md.alist[k][s]].d;
  w.acc[k] = Dual(pf.a[k], ad);
}

// primal f_D with the factor kept (x_next is written to xn if non-null)
H1_DEV void dyn_primal_factor_warp(const DynModel& md, DynWarp& w, const double* x, const double* u, double* xn,
                                   PrimalFactor& pf) {
  H1_LANES_DECL(double)
  H1_PHASE((ph_load<double, DynWarp>(lane, md, w, x, u, -1, 0.0)))
  H1_PHASE(ph_walk<double>(lane, md, w, H1_LR))
  H1_PHASE(ph_sums<double>(lane, md, w, H1_LR))
  H1_PHASE(ph_rows<double>(lane, md, w, H1_LR))
  for (int k = NV - 1; k >= 1; --k) H1_PHASE(ph_ltdl(lane, k, md, w))
  for (int lev = 0; lev < MAXSLOT; ++lev) H1_PHASE(ph_back(lane, lev, md, w))
  H1_PHASE(ph_save_factor(lane, md, w, pf))
  if (xn) H1_PHASE((ph_integrate<double, DynWarp>(lane, md, w, xn)))
}

// assembly of one tangent direction (no dependence on the primal factor)
H1_DEV void dyn_tangent_assemble_warp(const DynModel& md, DynWarpT<Dual>& w, const double* x, const double* u,
                                      int seed) {
  H1_LANES_DECL(Dual)
  H1_PHASE((ph_load<Dual, DynWarpT<Dual> >(lane, md, w, x, u, seed, 0.0)))
  H1_PHASE(ph_walk<Dual>(lane, md, w, H1_LR))
  H1_PHASE(ph_sums<Dual>(lane, md, w, H1_LR))
  H1_PHASE(ph_rows<Dual>(lane, md, w, H1_LR))
}
// tangent solve + integration; writes d x_next / d(input seed) (51 entries) to out_col
H1_DEV void dyn_tangent_solve_warp(const DynModel& md, DynWarpT<Dual>& w, const PrimalFactor& pf, double* out_col) {
#if defined(__CUDACC__)
  const int lane = threadIdx.x & 31;
#endif
  H1_PHASE(ph_tan_rhs(lane, md, w, pf))
  for (int lev = MAXSLOT - 1; lev >= 0; --lev) H1_PHASE(ph_tan_fwd(lane, lev, md, w, pf))
  for (int lev = 0; lev < MAXSLOT; ++lev) H1_PHASE(ph_tan_back(lane, lev, md, w, pf))
  H1_PHASE((ph_integrate<Dual, DynWarpT<Dual> >(lane, md, w, out_col)))
}

// ------------------------------------------------------------------------------------------------------
// Inverse-dynamics form of the tangent right-hand side. With a (the primal acceleration) held fixed,
//     g(q, v, u; a) = Mhat(q) a - rhs(q, v, u) = ID(q, v, a) + (armature + h D) a + D v - tau(u) - sum_i J_i^T F_i ,
//     F_i = phi_i(q, v) - W_i(q) (J_i a)            (linearly-implicit contact force at sole point i)
// so that  t = rhsdot - Mhatdot a = -d g / d(direction): ONE recursive Newton-Euler pass on dual numbers per
// direction, no mass-matrix rows, no contact Jacobian table. Scratch per warp drops to ~8.7 KB.
// ------------------------------------------------------------------------------------------------------
template <class T> struct TanWarpT {
  T q[NQ], v[NV], tau[NV];
  T sn[NB], cs[NB];
  T S[NV][6];
  T body[NB][6];             // body wrench about the base origin [n; l]
  T part[4][6];
  T footR[H1_NFOOT][9], footr[H1_NFOOT][3], footV[H1_NFOOT][6], footVa[H1_NFOOT][6];
  T cw[NCPT][6];             // contact wrench of each sole point about the base origin
  T acc[NV];
  double tvec[NV];
};

template <class T> struct TanLaneT {
  T R[9], r[3], V[6], Va[6], Ab[6], S[6];
};

template <class T> H1_DEV void ph_id_walk(int lane, const DynModel& md, TanWarpT<T>& w, TanLaneT<T>& L,
                                          const double* a) {
  if (lane >= NV) return;
  const int b = lane < 6 ? 0 : lane - 5;
  T qn[4];
  quat_normalize(&w.q[3], qn);
  quat_to_mat(qn, L.R);
  L.r[0] = L.r[1] = L.r[2] = T(0.0);
  const T wb[3] = {w.v[3], w.v[4], w.v[5]};
  for (int i = 0; i < 3; ++i) {
    L.V[i] = L.R[3 * i] * wb[0] + L.R[3 * i + 1] * wb[1] + L.R[3 * i + 2] * wb[2];
    L.Va[i] = L.R[3 * i] * a[3] + L.R[3 * i + 1] * a[4] + L.R[3 * i + 2] * a[5];
    L.V[3 + i] = w.v[i];
    L.Va[3 + i] = T(a[i]);
  }
  {
    T vxw[3];
    cross3(L.V + 3, L.V, vxw);
    L.Ab[0] = L.Ab[1] = L.Ab[2] = T(0.0);
    L.Ab[3] = vxw[0] - md.gravity[0]; L.Ab[4] = vxw[1] - md.gravity[1]; L.Ab[5] = vxw[2] - md.gravity[2];
  }
  for (int i = 0; i < 6; ++i) L.S[i] = T(0.0);
  if (lane < 3) { L.S[3] = T(lane == 0 ? 1.0 : 0.0); L.S[4] = T(lane == 1 ? 1.0 : 0.0); L.S[5] = T(lane == 2 ? 1.0 : 0.0); }
  else if (lane < 6) { col_of(L.R, lane - 3, L.S); }
  const int dep = md.depth[b];
#pragma unroll 1
  for (int d = 1; d <= 5; ++d) {
    if (d <= dep) {
      const int an = md.anc_body[b][d];
      const double* p = md.pos[an];
      L.r[0] += L.R[0] * p[0] + L.R[1] * p[1] + L.R[2] * p[2];
      L.r[1] += L.R[3] * p[0] + L.R[4] * p[1] + L.R[5] * p[2];
      L.r[2] += L.R[6] * p[0] + L.R[7] * p[1] + L.R[8] * p[2];
      if (md.has_rfix[an]) {
        const double* F = md.rfix[an];
        T Tm[9];
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k)
            Tm[3 * i + k] = L.R[3 * i] * F[k] + L.R[3 * i + 1] * F[3 + k] + L.R[3 * i + 2] * F[6 + k];
        for (int i = 0; i < 9; ++i) L.R[i] = Tm[i];
      }
      const int ax = md.axis[an];
      rot_right(L.R, ax, w.sn[an], w.cs[an]);
      col_of(L.R, ax, L.S);
      cross3(L.r, L.S, L.S + 3);
      const T vj = w.v[5 + an];
      const double aj = a[5 + an];
      for (int i = 0; i < 6; ++i) { L.V[i] += L.S[i] * vj; L.Va[i] += L.S[i] * aj; }
      T c1[3], c2[3], c3[3];
      cross3(L.V, L.S, c1); cross3(L.V, L.S + 3, c2); cross3(L.V + 3, L.S, c3);
      L.Ab[0] += c1[0] * vj; L.Ab[1] += c1[1] * vj; L.Ab[2] += c1[2] * vj;
      L.Ab[3] += (c2[0] + c3[0]) * vj; L.Ab[4] += (c2[1] + c3[1]) * vj; L.Ab[5] += (c2[2] + c3[2]) * vj;
    }
  }
  for (int i = 0; i < 6; ++i) w.S[lane][i] = L.S[i];
  if (lane < 5) return;
  for (int f = 0; f < H1_NFOOT; ++f)
    if (b == md.foot_body[f]) {
      for (int i = 0; i < 9; ++i) w.footR[f][i] = L.R[i];
      for (int i = 0; i < 3; ++i) w.footr[f][i] = L.r[i];
      for (int i = 0; i < 6; ++i) { w.footV[f][i] = L.V[i]; w.footVa[f][i] = L.Va[i]; }
    }
  T I[10];
  {
    const double* ip = md.ipos[b];
    T c[3] = {L.r[0] + L.R[0] * ip[0] + L.R[1] * ip[1] + L.R[2] * ip[2],
              L.r[1] + L.R[3] * ip[0] + L.R[4] * ip[1] + L.R[5] * ip[2],
              L.r[2] + L.R[6] * ip[0] + L.R[7] * ip[1] + L.R[8] * ip[2]};
    const double* J = md.inertia[b];
    T Tm[9];
    for (int i = 0; i < 3; ++i) {
      Tm[3 * i + 0] = L.R[3 * i] * J[0] + L.R[3 * i + 1] * J[3] + L.R[3 * i + 2] * J[4];
      Tm[3 * i + 1] = L.R[3 * i] * J[3] + L.R[3 * i + 1] * J[1] + L.R[3 * i + 2] * J[5];
      Tm[3 * i + 2] = L.R[3 * i] * J[4] + L.R[3 * i + 1] * J[5] + L.R[3 * i + 2] * J[2];
    }
    const double m = md.mass[b];
    const T cc = dot3(c, c);
    I[0] = T(m); I[1] = m * c[0]; I[2] = m * c[1]; I[3] = m * c[2];
    I[4] = dot3(Tm, L.R) + m * (cc - c[0] * c[0]);
    I[5] = dot3(Tm + 3, L.R + 3) + m * (cc - c[1] * c[1]);
    I[6] = dot3(Tm + 6, L.R + 6) + m * (cc - c[2] * c[2]);
    I[7] = dot3(Tm, L.R + 3) - m * (c[0] * c[1]);
    I[8] = dot3(Tm, L.R + 6) - m * (c[0] * c[2]);
    I[9] = dot3(Tm + 3, L.R + 6) - m * (c[1] * c[2]);
  }
  T At[6], Ia[6], Iv[6];
  for (int i = 0; i < 6; ++i) At[i] = L.Va[i] + L.Ab[i];
  spi_apply(I, At, Ia);
  spi_apply(I, L.V, Iv);
  T a1[3], a2[3], a3[3];
  cross3(L.V, Iv, a1); cross3(L.V + 3, Iv + 3, a2); cross3(L.V, Iv + 3, a3);
  w.body[b][0] = Ia[0] + a1[0] + a2[0]; w.body[b][1] = Ia[1] + a1[1] + a2[1]; w.body[b][2] = Ia[2] + a1[2] + a2[2];
  w.body[b][3] = Ia[3] + a3[0]; w.body[b][4] = Ia[4] + a3[1]; w.body[b][5] = Ia[5] + a3[2];
}

template <class T> H1_DEV void ph_id_contact(int lane, const DynModel& md, TanWarpT<T>& w) {
  if (lane >= NCPT) return;
  const int f = lane / H1_NCP;
  const T* R = w.footR[f];
  const double* pt = md.foot_pts[lane];
  T rho[3] = {w.footr[f][0] + R[0] * pt[0] + R[1] * pt[1] + R[2] * pt[2],
              w.footr[f][1] + R[3] * pt[0] + R[4] * pt[1] + R[5] * pt[2],
              w.footr[f][2] + R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]};
  T t1[3], t2[3];
  cross3(w.footV[f], rho, t1);
  cross3(w.footVa[f], rho, t2);
  const T pd[3] = {w.footV[f][3] + t1[0], w.footV[f][4] + t1[1], w.footV[f][5] + t1[2]};
  const T pa[3] = {w.footVa[f][3] + t2[0], w.footVa[f][4] + t2[1], w.footVa[f][5] + t2[2]};
  const double h = md.h;
  const T dd = -(w.q[2] + rho[2]);
  const T root = sqrt_t(dd * dd + md.eps * md.eps);
  const T sp = 0.5 * (dd + root), al = 0.5 * (1.0 + dd / root);
  T F[3];
  F[0] = -(al * md.bt) * (pd[0] + h * pa[0]);
  F[1] = -(al * md.bt) * (pd[1] + h * pa[1]);
  F[2] = md.kn * sp - al * ((md.bn + h * md.kn) * pd[2] + (h * md.bn + h * h * md.kn) * pa[2]);
  T n[3];
  cross3(rho, F, n);
  w.cw[lane][0] = n[0]; w.cw[lane][1] = n[1]; w.cw[lane][2] = n[2];
  w.cw[lane][3] = F[0]; w.cw[lane][4] = F[1]; w.cw[lane][5] = F[2];
}

template <class T> H1_DEV void ph_id_sums(int lane, const DynModel& md, TanWarpT<T>& w, T* T6) {
  if (lane < 6 || lane >= NV) return;
  const int b = lane - 5, e = md.chain_end[b];
  for (int i = 0; i < 6; ++i) T6[i] = T(0.0);
#pragma unroll 1
  for (int k = b; k <= e; ++k) {
    for (int i = 0; i < 6; ++i) T6[i] += w.body[k][i];
    for (int f = 0; f < H1_NFOOT; ++f)
      if (k == md.foot_body[f])
        for (int c = 0; c < H1_NCP; ++c)
          for (int i = 0; i < 6; ++i) T6[i] -= w.cw[f * H1_NCP + c][i];
  }
  if (md.parent[b] == 0) {
    const int slot = md.base_child_slot[b];
    for (int i = 0; i < 6; ++i) w.part[slot][i] = T6[i];
  }
}

template <class T> H1_DEV void ph_id_resid(int lane, const DynModel& md, TanWarpT<T>& w, const TanLaneT<T>& L, T* T6,
                                           const double* a) {
  if (lane >= NV) return;
  if (lane < 6) {
    for (int i = 0; i < 6; ++i) T6[i] = w.body[0][i];
    for (int c = 0; c < md.n_base_children; ++c)
      for (int i = 0; i < 6; ++i) T6[i] += w.part[c][i];
  }
  const int j = lane;
  const T g = dot6(L.S, T6) + (md.armature[j] + md.h * md.damping[j]) * a[j] + md.damping[j] * w.v[j] - w.tau[j];
  w.tvec[j] = -tangent_of(g);
  w.acc[j] = T(a[j]);
}
template <class W> H1_DEV void ph_ctrl_rhs(int lane, W& w, const double* a) {  // control direction: t = d tau
  if (lane >= NV) return;
  w.tvec[lane] = w.tau[lane].d;
  w.acc[lane] = Dual(a[lane], 0.0);
}

template <class W> H1_DEV void ph_tan_fwd_g(int lane, int lev, const DynModel& md, W& w, const PrimalFactor& pf) {
  if (lane >= NV || md.level[lane] != lev) return;
  const int i = lane, slot = md.nlist[i] - 1, e = md.dof_sub_end[i];
  double y = w.tvec[i];
  for (int k = i + 1; k <= e; ++k) y -= pf.Lm[k][slot] * w.tvec[k];
  w.tvec[i] = y;
}
template <class W> H1_DEV void ph_tan_back_g(int lane, int lev, const DynModel& md, W& w, const PrimalFactor& pf) {
  if (lane >= NV || md.level[lane] != lev) return;
  const int k = lane, n = md.nlist[k];
  double ad = w.tvec[k] / pf.D[k];
  for (int s = 0; s < n - 1; ++s) ad -= pf.Lm[k][s] * w.acc[md.alist[k][s]].d;
  w.acc[k] = Dual(pf.a[k], ad);
}

#if defined(__CUDACC__)
#define H1_TLANES_DECL TanLaneT<Dual> Lr; Dual T6r[6]; const int lane = threadIdx.x & 31;
#define H1_TLR Lr
#define H1_T6 T6r
#else
#define H1_TLANES_DECL static thread_local TanLaneT<Dual> Lr_[32]; static thread_local Dual T6r_[32][6];
#define H1_TLR Lr_[lane]
#define H1_T6 T6r_[lane]
#endif

// One exact column of [A | B]: d f_D / d(input `seed`), written to out_col[51].
H1_DEV void dyn_tangent_id_warp(const DynModel& md, TanWarpT<Dual>& w, const PrimalFactor& pf, const double* x,
                                const double* u, int seed, double* out_col) {
  H1_TLANES_DECL
  H1_PHASE((ph_load<Dual, TanWarpT<Dual> >(lane, md, w, x, u, seed, 0.0)))
  if (seed >= NX) {
    H1_PHASE(ph_ctrl_rhs(lane, w, pf.a))
  } else {
    H1_PHASE(ph_id_walk<Dual>(lane, md, w, H1_TLR, pf.a))
    H1_PHASE(ph_id_contact<Dual>(lane, md, w))
    H1_PHASE(ph_id_sums<Dual>(lane, md, w, H1_T6))
    H1_PHASE(ph_id_resid<Dual>(lane, md, w, H1_TLR, H1_T6, pf.a))
  }
  for (int lev = MAXSLOT - 1; lev >= 0; --lev) H1_PHASE(ph_tan_fwd_g(lane, lev, md, w, pf))
  for (int lev = 0; lev < MAXSLOT; ++lev) H1_PHASE(ph_tan_back_g(lane, lev, md, w, pf))
  H1_PHASE((ph_integrate<Dual, TanWarpT<Dual> >(lane, md, w, out_col)))
}

}  // namespace h1
