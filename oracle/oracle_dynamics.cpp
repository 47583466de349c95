// TEST INFRASTRUCTURE ONLY — CPU oracle of the one-step dynamics map f_D and its forward-difference
// linearization. Parity status: UNPINNED against MuJoCo (the reference's dynamics engine is an
// un-vendored third-party library, absent here; see DESIGN.md "Dynamics definition f_D").
//
// Structure follows the reference call sites:
//   RobotUtils::rolloutOneStep      /root/reference/src/common/robot_utils.cpp:106-117
//     (state in -> one discrete step of length timestep -> state out, raw coordinates)
//   RobotUtils::linearizeDynamicsFD /root/reference/src/common/robot_utils.cpp:120-160
//     (forward differences, eps = 1e-5 (include/common/robot_utils.hpp:51-53), one column per raw
//      state / control coordinate, 1 + 51 + 19 evaluations)
//   RobotUtils::computeCoM          /root/reference/src/common/robot_utils.cpp:810-833
//   RobotUtils::computeGravComp     /root/reference/src/common/robot_utils.cpp:844-866
// The arithmetic of the step itself is the documented map f_D, a textbook
// body-by-body spatial-algebra implementation: dense mass matrix, dense Cholesky.
#include "oracle.hpp"
#include "oracle_ad.hpp"
#include "oracle_math.hpp"
#include <cstring>
#include <vector>

namespace orc {

namespace {

template <class T> struct SpI {  // spatial inertia about the reference point, world-aligned axes
  T m, h[3], I[6];  // I: xx yy zz xy xz yz
};

template <class T> inline void spi_apply(const SpI<T>& s, const T* V, T* P) {
  // motion V=[w;v] -> momentum P=[n;l]:  n = I w + h x v ,  l = m v - h x w
  const T* w = V; const T* v = V + 3;
  T hv[3], hw[3];
  cross3(s.h, v, hv); cross3(s.h, w, hw);
  P[0] = s.I[0] * w[0] + s.I[3] * w[1] + s.I[4] * w[2] + hv[0];
  P[1] = s.I[3] * w[0] + s.I[1] * w[1] + s.I[5] * w[2] + hv[1];
  P[2] = s.I[4] * w[0] + s.I[5] * w[1] + s.I[2] * w[2] + hv[2];
  P[3] = s.m * v[0] - hw[0];
  P[4] = s.m * v[1] - hw[1];
  P[5] = s.m * v[2] - hw[2];
}
template <class T> inline void motion_cross(const T* V, const T* S, T* o) {  // V x S
  T a[3], b[3], c[3];
  cross3(V, S, a); cross3(V, S + 3, b); cross3(V + 3, S, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
template <class T> inline void force_cross(const T* V, const T* P, T* o) {  // V x* P
  T a[3], b[3], c[3];
  cross3(V, P, a); cross3(V + 3, P + 3, b); cross3(V, P + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
template <class T> inline T dot6(const T* a, const T* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// unit quaternion (w,x,y,z) -> rotation, MuJoCo-style homogeneous form
template <class T> inline void quat_to_mat(const T* q, T* R) {
  T q00 = q[0] * q[0], q11 = q[1] * q[1], q22 = q[2] * q[2], q33 = q[3] * q[3];
  T q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q12 = q[1] * q[2], q13 = q[1] * q[3], q23 = q[2] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[1] = 2.0 * (q12 - q03);     R[2] = 2.0 * (q13 + q02);
  R[3] = 2.0 * (q12 + q03);     R[4] = q00 - q11 + q22 - q33; R[5] = 2.0 * (q23 - q01);
  R[6] = 2.0 * (q13 - q02);     R[7] = 2.0 * (q23 + q01);     R[8] = q00 - q11 - q22 + q33;
}

template <class T> inline void normalize_quat_t(const T* q, T* qn) {
  T n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-12) { qn[0] = T(1.0); qn[1] = qn[2] = qn[3] = T(0.0); return; }
  for (int i = 0; i < 4; ++i) qn[i] = q[i] / n;
}

// Forward kinematics of the dynamics model: rotation R[b] (world-aligned) and origin r[b]
// RELATIVE TO THE BASE ORIGIN. The base quaternion is normalised first.
template <class T> void dyn_fk_t(const H1Model& md, const T* q, T (*R)[9], T (*r)[3]) {
  T qn[4];
  normalize_quat_t(q + 3, qn);
  quat_to_mat(qn, R[0]);
  r[0][0] = r[0][1] = r[0][2] = T(0.0);
  for (int b = 1; b < H1_NB; ++b) {
    int p = md.parent[b];
    T off[3];
    matvec3(R[p], md.pos[b], off);
    for (int i = 0; i < 3; ++i) r[b][i] = r[p][i] + off[i];
    if (md.has_rfix[b]) matmul3(R[p], md.rfix[b], R[b]);
    else for (int i = 0; i < 9; ++i) R[b][i] = R[p][i];
    T th = q[6 + b];
    rot_axis_right(R[b], md.axis[b], sin(th), cos(th));
  }
}

}  // namespace

void normalize_quat(const double* q, double* qn) { normalize_quat_t<double>(q, qn); }
void dyn_fk(const H1Model& md, const double* q, double (*R)[9], double (*r)[3]) { dyn_fk_t<double>(md, q, R, r); }

void dyn_com(const H1Model& md, const double* x, double* com) {
  double R[H1_NB][9], r[H1_NB][3];
  dyn_fk(md, x, R, r);
  double tot = 0, acc[3] = {0, 0, 0};
  for (int b = 0; b < H1_NB; ++b) {
    double c[3];
    matvec3(R[b], md.ipos[b], c);
    tot += md.mass[b];
    for (int i = 0; i < 3; ++i) acc[i] += md.mass[b] * (x[i] + r[b][i] + c[i]);
  }
  for (int i = 0; i < 3; ++i) com[i] = acc[i] / tot;
}

// Whole-body CoM velocity in the world frame: the value of mj_jacSubtreeCom(root) * qvel that
// RobotUtils::loadReferences stores per reference row (/root/reference/src/common/robot_utils.cpp:388-397).
// Body-by-body velocity propagation (origin velocity + angular velocity per body), mass-weighted mean of the
// body-CoM velocities. qvel conventions: world-frame base linear velocity, body-frame base angular velocity.
void dyn_com_vel(const H1Model& md, const double* x, double* cv) {
  double R[H1_NB][9], r[H1_NB][3];
  dyn_fk(md, x, R, r);
  double w[H1_NB][3], vo[H1_NB][3];
  const double* v = x + H1_NQ;
  matvec3(R[0], v + 3, w[0]);
  for (int i = 0; i < 3; ++i) vo[0][i] = v[i];
  double tot = 0, acc[3] = {0, 0, 0};
  for (int b = 0; b < H1_NB; ++b) {
    if (b > 0) {
      const int p = md.parent[b], ax = md.axis[b];
      double off[3], t[3];
      for (int i = 0; i < 3; ++i) off[i] = r[b][i] - r[p][i];
      cross3(w[p], off, t);
      for (int i = 0; i < 3; ++i) {
        vo[b][i] = vo[p][i] + t[i];
        w[b][i] = w[p][i] + R[b][3 * i + ax] * v[5 + b];
      }
    }
    double c[3], t[3];
    matvec3(R[b], md.ipos[b], c);
    cross3(w[b], c, t);
    tot += md.mass[b];
    for (int i = 0; i < 3; ++i) acc[i] += md.mass[b] * (vo[b][i] + t[i]);
  }
  for (int i = 0; i < 3; ++i) cv[i] = acc[i] / tot;
}

void dyn_body_pos(const H1Model& md, const double* x, int body, double* p) {
  double R[H1_NB][9], r[H1_NB][3];
  dyn_fk(md, x, R, r);
  for (int i = 0; i < 3; ++i) p[i] = x[i] + r[body][i];
}

// Core of f_D: assemble Mhat (dense, row-major [25][25]) and rhs at (q,v,u); optionally the plain bias forces.
template <class T>
static void assemble(const H1Model& md, const T* x, const T* u, T* Mh, T* rhs, T* bias_out) {
  const T* q = x; const T* v = x + H1_NQ;
  const double h = md.timestep;
  std::vector<T> Rv(H1_NB * 9), rv(H1_NB * 3);
  T (*R)[9] = reinterpret_cast<T(*)[9]>(Rv.data());
  T (*r)[3] = reinterpret_cast<T(*)[3]>(rv.data());
  dyn_fk_t<T>(md, q, R, r);

  // motion subspaces about the base origin, world-aligned: S = [angular; linear]
  std::vector<T> Sv(H1_NV * 6, T(0.0));
  T (*S)[6] = reinterpret_cast<T(*)[6]>(Sv.data());
  for (int k = 0; k < 3; ++k) S[k][3 + k] = T(1.0);
  for (int k = 0; k < 3; ++k) { S[3 + k][0] = R[0][k]; S[3 + k][1] = R[0][3 + k]; S[3 + k][2] = R[0][6 + k]; }
  for (int b = 1; b < H1_NB; ++b) {
    int j = 5 + b, ax = md.axis[b];
    T a[3] = {R[b][ax], R[b][3 + ax], R[b][6 + ax]};
    T m0[3];
    cross3(r[b], a, m0);
    for (int i = 0; i < 3; ++i) { S[j][i] = a[i]; S[j][3 + i] = m0[i]; }
  }
  // spatial inertias about the base origin
  std::vector<SpI<T>> I(H1_NB);
  for (int b = 0; b < H1_NB; ++b) {
    T c[3];
    matvec3(R[b], md.ipos[b], c);
    for (int i = 0; i < 3; ++i) c[i] = c[i] + r[b][i];
    const double* J = md.inertia[b];
    double Jb[9] = {J[0], J[3], J[4], J[3], J[1], J[5], J[4], J[5], J[2]};
    T Tm[9], Rt[9], Jw[9];
    matmul3(R[b], Jb, Tm);
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) Rt[3 * i + k] = R[b][3 * k + i];
    matmul3(Tm, Rt, Jw);
    double m = md.mass[b];
    T cc = dot3(c, c);
    I[b].m = T(m);
    for (int i = 0; i < 3; ++i) I[b].h[i] = m * c[i];
    I[b].I[0] = Jw[0] + m * (cc - c[0] * c[0]);
    I[b].I[1] = Jw[4] + m * (cc - c[1] * c[1]);
    I[b].I[2] = Jw[8] + m * (cc - c[2] * c[2]);
    I[b].I[3] = Jw[1] - m * (c[0] * c[1]);
    I[b].I[4] = Jw[2] - m * (c[0] * c[2]);
    I[b].I[5] = Jw[5] - m * (c[1] * c[2]);
  }
  // velocities and bias accelerations (qacc = 0)
  std::vector<T> Vv(H1_NB * 6, T(0.0)), Abv(H1_NB * 6, T(0.0)), Fv(H1_NB * 6, T(0.0));
  T (*V)[6] = reinterpret_cast<T(*)[6]>(Vv.data());
  T (*Ab)[6] = reinterpret_cast<T(*)[6]>(Abv.data());
  T (*F)[6] = reinterpret_cast<T(*)[6]>(Fv.data());
  for (int k = 0; k < 6; ++k) for (int i = 0; i < 6; ++i) V[0][i] = V[0][i] + S[k][i] * v[k];
  for (int i = 0; i < 3; ++i) Ab[0][3 + i] = T(-md.gravity[i]);
  for (int k = 3; k < 6; ++k) {
    T Sd[6];
    motion_cross(V[0], S[k], Sd);
    for (int i = 0; i < 6; ++i) Ab[0][i] = Ab[0][i] + Sd[i] * v[k];
  }
  for (int b = 1; b < H1_NB; ++b) {
    int p = md.parent[b], j = 5 + b;
    for (int i = 0; i < 6; ++i) V[b][i] = V[p][i] + S[j][i] * v[j];
    T Sd[6];
    motion_cross(V[b], S[j], Sd);
    for (int i = 0; i < 6; ++i) Ab[b][i] = Ab[p][i] + Sd[i] * v[j];
  }
  for (int b = 0; b < H1_NB; ++b) {
    T Ia[6], Iv[6], vx[6];
    spi_apply(I[b], Ab[b], Ia);
    spi_apply(I[b], V[b], Iv);
    force_cross(V[b], Iv, vx);
    for (int i = 0; i < 6; ++i) F[b][i] = Ia[i] + vx[i];
  }
  // subtree accumulation (children have larger indices than parents)
  std::vector<SpI<T>> Ic(I);
  for (int b = H1_NB - 1; b >= 1; --b) {
    int p = md.parent[b];
    for (int i = 0; i < 6; ++i) F[p][i] = F[p][i] + F[b][i];
    Ic[p].m = Ic[p].m + Ic[b].m;
    for (int i = 0; i < 3; ++i) Ic[p].h[i] = Ic[p].h[i] + Ic[b].h[i];
    for (int i = 0; i < 6; ++i) Ic[p].I[i] = Ic[p].I[i] + Ic[b].I[i];
  }
  std::vector<T> bias(H1_NV);
  for (int j = 0; j < H1_NV; ++j) bias[j] = dot6(S[j], F[j < 6 ? 0 : j - 5]);
  if (bias_out) for (int j = 0; j < H1_NV; ++j) bias_out[j] = bias[j];

  // composite-rigid-body mass matrix
  auto M = [&](int j, int k) -> T& { return Mh[j * H1_NV + k]; };
  for (int j = 0; j < H1_NV * H1_NV; ++j) Mh[j] = T(0.0);
  for (int j = 0; j < H1_NV; ++j) {
    int b = j < 6 ? 0 : j - 5;
    T P[6];
    spi_apply(Ic[b], S[j], P);
    if (j < 6) {
      for (int k = 0; k <= j; ++k) { T val = dot6(S[k], P); M(j, k) = val; M(k, j) = val; }
    } else {
      M(j, j) = dot6(S[j], P);
      for (int a = md.parent[b]; a >= 0; a = md.parent[a]) {
        if (a == 0) { for (int k = 0; k < 6; ++k) { T val = dot6(S[k], P); M(j, k) = val; M(k, j) = val; } }
        else { int k = 5 + a; T val = dot6(S[k], P); M(j, k) = val; M(k, j) = val; }
      }
    }
  }
  for (int j = 0; j < H1_NV; ++j) {
    M(j, j) = M(j, j) + (md.armature[j] + h * md.damping[j]);
    T tau(0.0);
    if (j >= 6 && u) {
      tau = u[j - 6];
      if (tau < md.ctrl_range[j - 6][0]) tau = T(md.ctrl_range[j - 6][0]);
      if (tau > md.ctrl_range[j - 6][1]) tau = T(md.ctrl_range[j - 6][1]);
    }
    rhs[j] = tau - bias[j] - md.damping[j] * v[j];
  }
  // soft sole contacts, linearly-implicit in the point velocity and height
  std::vector<T> Jv(3 * H1_NV);
  for (int f = 0; f < H1_NFOOT; ++f) {
    int fb = md.foot_body[f];
    for (int c = 0; c < H1_NCP; ++c) {
      T rho[3];
      matvec3(R[fb], md.foot_pts[f][c], rho);
      for (int i = 0; i < 3; ++i) rho[i] = rho[i] + r[fb][i];
      for (auto& e : Jv) e = T(0.0);
      auto J = [&](int i, int k) -> T& { return Jv[i * H1_NV + k]; };
      auto col = [&](int k) {
        T t[3];
        cross3(S[k], rho, t);
        for (int i = 0; i < 3; ++i) J(i, k) = S[k][3 + i] + t[i];
      };
      for (int k = 0; k < 6; ++k) col(k);
      for (int a = fb; a >= 1; a = md.parent[a]) col(5 + a);
      T pd[3] = {T(0.0), T(0.0), T(0.0)};
      for (int k = 0; k < H1_NV; ++k) for (int i = 0; i < 3; ++i) pd[i] = pd[i] + J(i, k) * v[k];
      T d = -(q[2] + rho[2]);
      T root = sqrt(d * d + md.contact_eps * md.contact_eps);
      T sp = 0.5 * (d + root), al = 0.5 * (1.0 + d / root);
      T W[3] = {al * (h * md.contact_bt), al * (h * md.contact_bt), al * (h * md.contact_bn + h * h * md.contact_kn)};
      T phi[3] = {-(al * md.contact_bt) * pd[0], -(al * md.contact_bt) * pd[1],
                  md.contact_kn * sp - al * (md.contact_bn + h * md.contact_kn) * pd[2]};
      for (int j = 0; j < H1_NV; ++j) {
        rhs[j] = rhs[j] + J(0, j) * phi[0] + J(1, j) * phi[1] + J(2, j) * phi[2];
        for (int k = 0; k < H1_NV; ++k)
          M(j, k) = M(j, k) + W[0] * J(0, j) * J(0, k) + W[1] * J(1, j) * J(1, k) + W[2] * J(2, j) * J(2, k);
      }
    }
  }
}

void dyn_bias(const H1Model& md, const double* x, double* bias) {
  std::vector<double> Mh(H1_NV * H1_NV), rhs(H1_NV);
  assemble<double>(md, x, nullptr, Mh.data(), rhs.data(), bias);
}

template <class T> static bool chol_solve(T* A, T* b, int n) {  // A row-major n x n (lower used), in place
  for (int j = 0; j < n; ++j) {
    T d = A[j * n + j];
    for (int k = 0; k < j; ++k) d = d - A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      T s = A[i * n + j];
      for (int k = 0; k < j; ++k) s = s - A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { T s = b[i]; for (int k = 0; k < i; ++k) s = s - A[i * n + k] * b[k]; b[i] = s / A[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { T s = b[i]; for (int k = i + 1; k < n; ++k) s = s - A[k * n + i] * b[k]; b[i] = s / A[i * n + i]; }
  return true;
}

// x_next = f_D(x, u)
template <class T> static void dyn_step_t(const H1Model& md, const T* x, const T* u, T* xn) {
  std::vector<T> Mh(H1_NV * H1_NV), acc(H1_NV);
  assemble<T>(md, x, u, Mh.data(), acc.data(), nullptr);
  chol_solve<T>(Mh.data(), acc.data(), H1_NV);
  const double h = md.timestep;
  const T* q = x; const T* v = x + H1_NQ;
  T* qn = xn; T* vn = xn + H1_NQ;
  for (int j = 0; j < H1_NV; ++j) vn[j] = v[j] + h * acc[j];
  for (int i = 0; i < 3; ++i) qn[i] = q[i] + h * vn[i];
  T qu[4];
  normalize_quat_t(q + 3, qu);
  T ph[3] = {h * vn[3], h * vn[4], h * vn[5]};
  T ang = sqrt(ph[0] * ph[0] + ph[1] * ph[1] + ph[2] * ph[2]);
  T e[4];
  if (ang < 1e-10) { e[0] = T(1.0); e[1] = 0.5 * ph[0]; e[2] = 0.5 * ph[1]; e[3] = 0.5 * ph[2]; }
  else { T sc = sin(0.5 * ang) / ang; e[0] = cos(0.5 * ang); e[1] = sc * ph[0]; e[2] = sc * ph[1]; e[3] = sc * ph[2]; }
  T pq[4] = {qu[0] * e[0] - qu[1] * e[1] - qu[2] * e[2] - qu[3] * e[3],
             qu[0] * e[1] + qu[1] * e[0] + qu[2] * e[3] - qu[3] * e[2],
             qu[0] * e[2] - qu[1] * e[3] + qu[2] * e[0] + qu[3] * e[1],
             qu[0] * e[3] + qu[1] * e[2] - qu[2] * e[1] + qu[3] * e[0]};
  normalize_quat_t(pq, qn + 3);
  for (int b = 1; b < H1_NB; ++b) qn[6 + b] = q[6 + b] + h * vn[5 + b];
}

void dyn_step(const H1Model& md, const double* x, const double* u, double* xn) { dyn_step_t<double>(md, x, u, xn); }

// Forward-difference linearization, column-major A[51x51], B[51x19] (Eigen layout).
void dyn_linearize_fd(const H1Model& md, const double* x, const double* u, double eps, double* A, double* B) {
  double base[H1_NX], xp[H1_NX], up[H1_NU], fp[H1_NX];
  dyn_step(md, x, u, base);
  for (int i = 0; i < H1_NX; ++i) {
    std::memcpy(xp, x, sizeof(xp));
    xp[i] += eps;
    dyn_step(md, xp, u, fp);
    for (int r = 0; r < H1_NX; ++r) A[i * H1_NX + r] = (fp[r] - base[r]) / eps;
  }
  for (int j = 0; j < H1_NU; ++j) {
    std::memcpy(up, u, sizeof(up));
    up[j] += eps;
    dyn_step(md, x, up, fp);
    for (int r = 0; r < H1_NX; ++r) B[j * H1_NX + r] = (fp[r] - base[r]) / eps;
  }
}

// Exact linearization: forward-mode AD through f_D with one tangent per input coordinate (70 directions).
void dyn_linearize_ad(const H1Model& md, const double* x, const double* u, double* A, double* B) {
  typedef D1<H1_NX + H1_NU> AD;
  std::vector<AD> xa(H1_NX), ua(H1_NU), xn(H1_NX);
  for (int i = 0; i < H1_NX; ++i) xa[i] = AD::var(x[i], i);
  for (int j = 0; j < H1_NU; ++j) ua[j] = AD::var(u[j], H1_NX + j);
  dyn_step_t<AD>(md, xa.data(), ua.data(), xn.data());
  for (int i = 0; i < H1_NX; ++i) for (int r = 0; r < H1_NX; ++r) A[i * H1_NX + r] = xn[r].d[i];
  for (int j = 0; j < H1_NU; ++j) for (int r = 0; r < H1_NX; ++r) B[j * H1_NX + r] = xn[r].d[H1_NX + j];
}

}  // namespace orc
