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
#include "oracle_math.hpp"
#include <cstring>

namespace orc {

namespace {

struct SpI {  // spatial inertia about the reference point, world-aligned axes
  double m, h[3], I[6];  // I: xx yy zz xy xz yz
};

inline void spi_apply(const SpI& s, const double* V, double* P) {
  // motion V=[w;v] -> momentum P=[n;l]:  n = I w + h x v ,  l = m v - h x w
  const double* w = V; const double* v = V + 3;
  double hv[3], hw[3];
  cross3(s.h, v, hv); cross3(s.h, w, hw);
  P[0] = s.I[0] * w[0] + s.I[3] * w[1] + s.I[4] * w[2] + hv[0];
  P[1] = s.I[3] * w[0] + s.I[1] * w[1] + s.I[5] * w[2] + hv[1];
  P[2] = s.I[4] * w[0] + s.I[5] * w[1] + s.I[2] * w[2] + hv[2];
  P[3] = s.m * v[0] - hw[0];
  P[4] = s.m * v[1] - hw[1];
  P[5] = s.m * v[2] - hw[2];
}
inline void motion_cross(const double* V, const double* S, double* o) {  // V x S
  double a[3], b[3], c[3];
  cross3(V, S, a); cross3(V, S + 3, b); cross3(V + 3, S, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
inline void force_cross(const double* V, const double* P, double* o) {  // V x* P
  double a[3], b[3], c[3];
  cross3(V, P, a); cross3(V + 3, P + 3, b); cross3(V, P + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
inline double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// unit quaternion (w,x,y,z) -> rotation, MuJoCo-style homogeneous form
inline void quat_to_mat(const double* q, double* R) {
  double q00 = q[0] * q[0], q11 = q[1] * q[1], q22 = q[2] * q[2], q33 = q[3] * q[3];
  double q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q12 = q[1] * q[2], q13 = q[1] * q[3], q23 = q[2] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[1] = 2 * (q12 - q03);       R[2] = 2 * (q13 + q02);
  R[3] = 2 * (q12 + q03);       R[4] = q00 - q11 + q22 - q33; R[5] = 2 * (q23 - q01);
  R[6] = 2 * (q13 - q02);       R[7] = 2 * (q23 + q01);       R[8] = q00 - q11 - q22 + q33;
}

}  // namespace

void normalize_quat(const double* q, double* qn) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-12) { qn[0] = 1; qn[1] = qn[2] = qn[3] = 0; return; }
  for (int i = 0; i < 4; ++i) qn[i] = q[i] / n;
}

// Forward kinematics of the dynamics model: rotation R[b] (world-aligned) and origin r[b]
// RELATIVE TO THE BASE ORIGIN. The base quaternion is normalised first.
void dyn_fk(const H1Model& md, const double* q, double (*R)[9], double (*r)[3]) {
  double qn[4];
  normalize_quat(q + 3, qn);
  quat_to_mat(qn, R[0]);
  r[0][0] = r[0][1] = r[0][2] = 0.0;
  for (int b = 1; b < H1_NB; ++b) {
    int p = md.parent[b];
    double off[3];
    matvec3(R[p], md.pos[b], off);
    for (int i = 0; i < 3; ++i) r[b][i] = r[p][i] + off[i];
    if (md.has_rfix[b]) matmul3(R[p], md.rfix[b], R[b]);
    else std::memcpy(R[b], R[p], sizeof(double) * 9);
    double th = q[6 + b];
    rot_axis_right(R[b], md.axis[b], std::sin(th), std::cos(th));
  }
}

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

void dyn_body_pos(const H1Model& md, const double* x, int body, double* p) {
  double R[H1_NB][9], r[H1_NB][3];
  dyn_fk(md, x, R, r);
  for (int i = 0; i < 3; ++i) p[i] = x[i] + r[body][i];
}

// Core of f_D: assemble Mhat and rhs at (q,v,u); optionally return the plain bias forces.
static void assemble(const H1Model& md, const double* x, const double* u, double (*Mh)[H1_NV], double* rhs,
                     double* bias_out) {
  const double* q = x; const double* v = x + H1_NQ;
  const double h = md.timestep;
  double R[H1_NB][9], r[H1_NB][3];
  dyn_fk(md, q, R, r);

  // motion subspaces about the base origin, world-aligned: S = [angular; linear]
  double S[H1_NV][6];
  std::memset(S, 0, sizeof(S));
  for (int k = 0; k < 3; ++k) S[k][3 + k] = 1.0;
  for (int k = 0; k < 3; ++k) { S[3 + k][0] = R[0][k]; S[3 + k][1] = R[0][3 + k]; S[3 + k][2] = R[0][6 + k]; }
  for (int b = 1; b < H1_NB; ++b) {
    int j = 5 + b, ax = md.axis[b];
    double a[3] = {R[b][ax], R[b][3 + ax], R[b][6 + ax]};
    double m0[3];
    cross3(r[b], a, m0);
    for (int i = 0; i < 3; ++i) { S[j][i] = a[i]; S[j][3 + i] = m0[i]; }
  }
  // spatial inertias about the base origin
  SpI I[H1_NB];
  for (int b = 0; b < H1_NB; ++b) {
    double c[3];
    matvec3(R[b], md.ipos[b], c);
    for (int i = 0; i < 3; ++i) c[i] += r[b][i];
    const double* J = md.inertia[b];
    double Jb[9] = {J[0], J[3], J[4], J[3], J[1], J[5], J[4], J[5], J[2]};
    double T[9], Rt[9], Jw[9];
    matmul3(R[b], Jb, T);
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) Rt[3 * i + k] = R[b][3 * k + i];
    matmul3(T, Rt, Jw);
    double m = md.mass[b], cc = dot3(c, c);
    I[b].m = m;
    for (int i = 0; i < 3; ++i) I[b].h[i] = m * c[i];
    I[b].I[0] = Jw[0] + m * (cc - c[0] * c[0]);
    I[b].I[1] = Jw[4] + m * (cc - c[1] * c[1]);
    I[b].I[2] = Jw[8] + m * (cc - c[2] * c[2]);
    I[b].I[3] = Jw[1] - m * c[0] * c[1];
    I[b].I[4] = Jw[2] - m * c[0] * c[2];
    I[b].I[5] = Jw[5] - m * c[1] * c[2];
  }
  // velocities and bias accelerations (qacc = 0)
  double V[H1_NB][6], Ab[H1_NB][6], F[H1_NB][6];
  for (int i = 0; i < 6; ++i) V[0][i] = 0;
  for (int k = 0; k < 6; ++k) for (int i = 0; i < 6; ++i) V[0][i] += S[k][i] * v[k];
  for (int i = 0; i < 3; ++i) { Ab[0][i] = 0; Ab[0][3 + i] = -md.gravity[i]; }
  for (int k = 3; k < 6; ++k) {
    double Sd[6];
    motion_cross(V[0], S[k], Sd);
    for (int i = 0; i < 6; ++i) Ab[0][i] += Sd[i] * v[k];
  }
  for (int b = 1; b < H1_NB; ++b) {
    int p = md.parent[b], j = 5 + b;
    for (int i = 0; i < 6; ++i) V[b][i] = V[p][i] + S[j][i] * v[j];
    double Sd[6];
    motion_cross(V[b], S[j], Sd);
    for (int i = 0; i < 6; ++i) Ab[b][i] = Ab[p][i] + Sd[i] * v[j];
  }
  for (int b = 0; b < H1_NB; ++b) {
    double Ia[6], Iv[6], vx[6];
    spi_apply(I[b], Ab[b], Ia);
    spi_apply(I[b], V[b], Iv);
    force_cross(V[b], Iv, vx);
    for (int i = 0; i < 6; ++i) F[b][i] = Ia[i] + vx[i];
  }
  // subtree accumulation (children have larger indices than parents)
  SpI Ic[H1_NB];
  for (int b = 0; b < H1_NB; ++b) Ic[b] = I[b];
  for (int b = H1_NB - 1; b >= 1; --b) {
    int p = md.parent[b];
    for (int i = 0; i < 6; ++i) F[p][i] += F[b][i];
    Ic[p].m += Ic[b].m;
    for (int i = 0; i < 3; ++i) Ic[p].h[i] += Ic[b].h[i];
    for (int i = 0; i < 6; ++i) Ic[p].I[i] += Ic[b].I[i];
  }
  double bias[H1_NV];
  for (int j = 0; j < H1_NV; ++j) bias[j] = dot6(S[j], F[j < 6 ? 0 : j - 5]);
  if (bias_out) std::memcpy(bias_out, bias, sizeof(bias));

  // composite-rigid-body mass matrix
  for (int j = 0; j < H1_NV; ++j) for (int k = 0; k < H1_NV; ++k) Mh[j][k] = 0.0;
  for (int j = 0; j < H1_NV; ++j) {
    int b = j < 6 ? 0 : j - 5;
    double P[6];
    spi_apply(Ic[b], S[j], P);
    if (j < 6) {
      for (int k = 0; k <= j; ++k) Mh[j][k] = Mh[k][j] = dot6(S[k], P);
    } else {
      Mh[j][j] = dot6(S[j], P);
      for (int a = md.parent[b]; a >= 0; a = md.parent[a]) {
        if (a == 0) { for (int k = 0; k < 6; ++k) Mh[j][k] = Mh[k][j] = dot6(S[k], P); }
        else { int k = 5 + a; Mh[j][k] = Mh[k][j] = dot6(S[k], P); }
      }
    }
  }
  for (int j = 0; j < H1_NV; ++j) {
    Mh[j][j] += md.armature[j] + h * md.damping[j];
    double tau = 0.0;
    if (j >= 6 && u) {
      tau = u[j - 6];
      if (tau < md.ctrl_range[j - 6][0]) tau = md.ctrl_range[j - 6][0];
      if (tau > md.ctrl_range[j - 6][1]) tau = md.ctrl_range[j - 6][1];
    }
    rhs[j] = tau - bias[j] - md.damping[j] * v[j];
  }
  // soft sole contacts, linearly-implicit in the point velocity and height
  for (int f = 0; f < H1_NFOOT; ++f) {
    int fb = md.foot_body[f];
    for (int c = 0; c < H1_NCP; ++c) {
      double rho[3];
      matvec3(R[fb], md.foot_pts[f][c], rho);
      for (int i = 0; i < 3; ++i) rho[i] += r[fb][i];
      double J[3][H1_NV];
      std::memset(J, 0, sizeof(J));
      auto col = [&](int k) {
        double t[3];
        cross3(S[k], rho, t);
        for (int i = 0; i < 3; ++i) J[i][k] = S[k][3 + i] + t[i];
      };
      for (int k = 0; k < 6; ++k) col(k);
      for (int a = fb; a >= 1; a = md.parent[a]) col(5 + a);
      double pd[3] = {0, 0, 0};
      for (int k = 0; k < H1_NV; ++k) for (int i = 0; i < 3; ++i) pd[i] += J[i][k] * v[k];
      double d = -(q[2] + rho[2]);
      double root = std::sqrt(d * d + md.contact_eps * md.contact_eps);
      double s = 0.5 * (d + root), al = 0.5 * (1.0 + d / root);
      double W[3] = {al * h * md.contact_bt, al * h * md.contact_bt,
                     al * (h * md.contact_bn + h * h * md.contact_kn)};
      double phi[3] = {-al * md.contact_bt * pd[0], -al * md.contact_bt * pd[1],
                       md.contact_kn * s - al * (md.contact_bn + h * md.contact_kn) * pd[2]};
      for (int j = 0; j < H1_NV; ++j) {
        rhs[j] += J[0][j] * phi[0] + J[1][j] * phi[1] + J[2][j] * phi[2];
        for (int k = 0; k < H1_NV; ++k)
          Mh[j][k] += W[0] * J[0][j] * J[0][k] + W[1] * J[1][j] * J[1][k] + W[2] * J[2][j] * J[2][k];
      }
    }
  }
}

void dyn_bias(const H1Model& md, const double* x, double* bias) {
  double Mh[H1_NV][H1_NV], rhs[H1_NV];
  assemble(md, x, nullptr, Mh, rhs, bias);
}

static bool chol_solve(double (*A)[H1_NV], double* b, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i][k] * b[k]; b[i] = s / A[i][i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[k][i] * b[k]; b[i] = s / A[i][i]; }
  return true;
}

// x_next = f_D(x, u)
void dyn_step(const H1Model& md, const double* x, const double* u, double* xn) {
  double Mh[H1_NV][H1_NV], acc[H1_NV];
  assemble(md, x, u, Mh, acc, nullptr);
  chol_solve(Mh, acc, H1_NV);
  const double h = md.timestep;
  const double* q = x; const double* v = x + H1_NQ;
  double* qn = xn; double* vn = xn + H1_NQ;
  for (int j = 0; j < H1_NV; ++j) vn[j] = v[j] + h * acc[j];
  for (int i = 0; i < 3; ++i) qn[i] = q[i] + h * vn[i];
  double qu[4];
  normalize_quat(q + 3, qu);
  double ph[3] = {h * vn[3], h * vn[4], h * vn[5]};
  double ang = std::sqrt(ph[0] * ph[0] + ph[1] * ph[1] + ph[2] * ph[2]);
  double e[4];
  if (ang < 1e-10) { e[0] = 1.0; e[1] = 0.5 * ph[0]; e[2] = 0.5 * ph[1]; e[3] = 0.5 * ph[2]; }
  else { double sc = std::sin(0.5 * ang) / ang; e[0] = std::cos(0.5 * ang); e[1] = sc * ph[0]; e[2] = sc * ph[1]; e[3] = sc * ph[2]; }
  double pq[4] = {qu[0] * e[0] - qu[1] * e[1] - qu[2] * e[2] - qu[3] * e[3],
                  qu[0] * e[1] + qu[1] * e[0] + qu[2] * e[3] - qu[3] * e[2],
                  qu[0] * e[2] - qu[1] * e[3] + qu[2] * e[0] + qu[3] * e[1],
                  qu[0] * e[3] + qu[1] * e[2] - qu[2] * e[1] + qu[3] * e[0]};
  normalize_quat(pq, qn + 3);
  for (int b = 1; b < H1_NB; ++b) qn[6 + b] = q[6 + b] + h * vn[5 + b];
}

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

}  // namespace orc
