// TEST INFRASTRUCTURE ONLY — CPU oracle of the kinematic cost terms and limit penalties.
// Parity status: UNPINNED against Pinocchio/CasADi (absent here); pinned instead against closed-form
// known answers and against finite differences of its own cost values.
//
// Restates:
//   derivatives::convertMuJoCoToPinocchio   /root/reference/src/common/derivatives.cpp:12-24
//   symDerivatives::symCoMPos / symCoMVel   derivatives.cpp:525-582
//   symDerivatives::symEEPos / symEEVel     derivatives.cpp:584-644
//   symDerivatives::symUpright / symBalance derivatives.cpp:646-707
//   (gradient / Hessian extraction)         derivatives.cpp:95-106,147-157,198-199,725-726; symmetrisation :521,:796
//   RobotUtils::constraintCost/Gradients/Hessians  /root/reference/src/common/robot_utils.cpp:615-778
// Pinocchio semantics restated here: free-flyer q~ = [p, qx,qy,qz,qw], rotation =
// Eigen quaternion polynomial WITHOUT normalisation, v~ = [v_lin (body frame), omega (body frame), joint
// rates]; forwardKinematics propagates local-frame spatial velocities; centerOfMass accumulates
// mass-weighted levers leaf->root; getFrameVelocity(LOCAL_WORLD_ALIGNED) = oRf * v_local.linear.
#include "oracle.hpp"
#include "oracle_ad.hpp"
#include "oracle_math.hpp"
#include <cstring>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Part 1: templated Pinocchio-style evaluation (runs on double for values, on D2<51> for derivatives)
// ---------------------------------------------------------------------------------------------
namespace {

template <class T> void quat_poly_rot(const T* xi, T* R) {  // xi = (x,y,z,w), Eigen toRotationMatrix polynomial
  const T &x = xi[0], &y = xi[1], &z = xi[2], &w = xi[3];
  T tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  T twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
    tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

template <class T> struct PinData {
  std::vector<T> oR, op, vw, vl, E;  // per body: 9, 3, 3, 3, 9
  PinData() : oR(H1_NB * 9), op(H1_NB * 3), vw(H1_NB * 3), vl(H1_NB * 3), E(H1_NB * 9) {}
};

// pinocchio::forwardKinematics(model, data, q, v)
template <class T> void pin_fk(const H1Model& cm, const T* qt, const T* vt, PinData<T>& d) {
  quat_poly_rot(qt + 3, &d.oR[0]);
  for (int i = 0; i < 3; ++i) { d.op[i] = qt[i]; d.vl[i] = vt[i]; d.vw[i] = vt[3 + i]; }
  for (int b = 1; b < H1_NB; ++b) {
    int p = cm.parent[b];
    T* E = &d.E[9 * b];
    for (int i = 0; i < 9; ++i) E[i] = T(cm.rfix[b][i]);
    T th = qt[6 + b];
    rot_axis_right(E, cm.axis[b], sin(th), cos(th));
    matmul3(&d.oR[9 * p], E, &d.oR[9 * b]);
    T t[3] = {T(cm.pos[b][0]), T(cm.pos[b][1]), T(cm.pos[b][2])}, off[3];
    matvec3(&d.oR[9 * p], t, off);
    for (int i = 0; i < 3; ++i) d.op[3 * b + i] = d.op[3 * p + i] + off[i];
    // v_i = S qdot + liMi.actInv(v_parent)
    T wxt[3], lin[3];
    cross3(&d.vw[3 * p], t, wxt);
    for (int i = 0; i < 3; ++i) lin[i] = d.vl[3 * p + i] + wxt[i];
    matTvec3(E, &d.vw[3 * p], &d.vw[3 * b]);
    matTvec3(E, lin, &d.vl[3 * b]);
    d.vw[3 * b + cm.axis[b]] = d.vw[3 * b + cm.axis[b]] + vt[5 + b];
  }
}

// pinocchio::centerOfMass(model, data, q, v): com[0], vcom[0]
template <class T> void pin_com(const H1Model& cm, const PinData<T>& d, T* com, T* vcom) {
  std::vector<T> cacc(H1_NB * 3), vacc(H1_NB * 3);
  std::vector<double> macc(H1_NB);
  for (int b = 0; b < H1_NB; ++b) {
    double m = cm.mass[b];
    T lever[3] = {T(cm.ipos[b][0]), T(cm.ipos[b][1]), T(cm.ipos[b][2])}, wl[3];
    cross3(&d.vw[3 * b], lever, wl);
    macc[b] = m;
    for (int i = 0; i < 3; ++i) { cacc[3 * b + i] = m * lever[i]; vacc[3 * b + i] = m * (wl[i] + d.vl[3 * b + i]); }
  }
  for (int b = H1_NB - 1; b >= 1; --b) {
    int p = cm.parent[b];
    T rc[3], rv[3];
    matvec3(&d.E[9 * b], &cacc[3 * b], rc);
    matvec3(&d.E[9 * b], &vacc[3 * b], rv);
    for (int i = 0; i < 3; ++i) {
      cacc[3 * p + i] = cacc[3 * p + i] + rc[i] + macc[b] * cm.pos[b][i];
      vacc[3 * p + i] = vacc[3 * p + i] + rv[i];
    }
    macc[p] += macc[b];
  }
  T rc[3], rv[3];
  matvec3(&d.oR[0], &cacc[0], rc);
  matvec3(&d.oR[0], &vacc[0], rv);
  for (int i = 0; i < 3; ++i) { com[i] = (rc[i] + macc[0] * d.op[i]) / macc[0]; vcom[i] = rv[i] / macc[0]; }
}

template <class T> T sq(const T& a) { return a * a; }

template <class T> T term_value(const H1Model& cm, int term, int ee, const T* xt, const double* target, double w) {
  const T* qt = xt; const T* vt = xt + H1_NQ;
  if (term == TERM_UPRIGHT) {  // reads x~[3..6] in the roles (qw,qx,qy,qz): reference quirk Q4
    const T &qw = xt[3], &qx = xt[4], &qy = xt[5], &qz = xt[6];
    T zx = 2.0 * (qx * qz + qw * qy), zy = 2.0 * (qy * qz - qw * qx), zz = 1.0 - 2.0 * (qx * qx + qy * qy);
    T rz = zz - 1.0;
    return (0.5 * w) * (zx * zx + zy * zy + rz * rz);
  }
  PinData<T> d;
  pin_fk(cm, qt, vt, d);
  if (term == TERM_EE_POS) {
    int b = cm.foot_body[ee];
    return w * (sq(d.op[3 * b] - target[0]) + sq(d.op[3 * b + 1] - target[1]) + sq(d.op[3 * b + 2] - target[2]));
  }
  if (term == TERM_EE_VEL) {
    int b = cm.foot_body[ee];
    T vel[3];
    matvec3(&d.oR[9 * b], &d.vl[3 * b], vel);
    return w * (sq(vel[0] - target[0]) + sq(vel[1] - target[1]) + sq(vel[2] - target[2]));
  }
  T com[3], vcom[3];
  pin_com(cm, d, com, vcom);
  if (term == TERM_COM) return w * (sq(com[0] - target[0]) + sq(com[1] - target[1]) + sq(com[2] - target[2]));
  if (term == TERM_COM_VEL) return w * (sq(vcom[0] - target[0]) + sq(vcom[1] - target[1]) + sq(vcom[2] - target[2]));
  // TERM_BALANCE: capture point, g = 9.81 hard-coded (quirk Q12)
  T omega0 = sqrt(com[2] / 9.81);
  T r0 = com[0] + vcom[0] * omega0 - target[0];
  T r1 = com[1] + vcom[1] * omega0 - target[1];
  return (0.5 * w) * (r0 * r0 + r1 * r1);
}

template <class T> void to_pin(const T* x_mj, T* xt) {  // convertMuJoCoToPinocchio: velocities untouched (Q5)
  for (int i = 0; i < H1_NX; ++i) xt[i] = x_mj[i];
  xt[3] = x_mj[4]; xt[4] = x_mj[5]; xt[5] = x_mj[6]; xt[6] = x_mj[3];
}

}  // namespace

double cost_term_value(const H1Model& cm, int term, int ee, const double* x_mj, const double* target, double w) {
  double xt[H1_NX];
  to_pin(x_mj, xt);
  return term_value<double>(cm, term, ee, xt, target, w);
}

void cost_term_ad(const H1Model& cm, int term, int ee, const double* x_mj, const double* target, double w, double* g,
                  double* H) {
  typedef D2<H1_NX> AD;
  double xt[H1_NX];
  to_pin(x_mj, xt);
  std::vector<AD> xa(H1_NX);
  for (int i = 0; i < H1_NX; ++i) xa[i] = AD::var(xt[i], i);
  AD c = term_value<AD>(cm, term, ee, xa.data(), target, w);
  for (int i = 0; i < H1_NX; ++i) g[i] += c.G(i);
  for (int j = 0; j < H1_NX; ++j)
    for (int i = 0; i < H1_NX; ++i) H[j * H1_NX + i] += c.H(i, j);
}

// ---------------------------------------------------------------------------------------------
// Part 2: hand-derived exact derivatives. Everything is expressed in the BASE frame:
//   P(z) = p_b + R(xi) r(theta)            (CoM or frame position)
//   U(z) = R(xi) u(theta, v),   u = v_b + w_b x r + sum_j r_j thdot_j     (CoM or frame velocity)
// where r is a weighted point set attached to the bodies (weights m_i/M for the CoM, an indicator for a
// frame), r_j = d r/d theta_j, etc. R is the un-normalised quaternion polynomial, so dR/dxi is linear in xi
// and d2R/dxi2 is constant.
// ---------------------------------------------------------------------------------------------
namespace {

struct BaseKin {
  double E[H1_NB][9], r[H1_NB][3], a[H1_NB][3], c[H1_NB][3];
  bool anc[H1_NB][H1_NB];  // anc[k][l]: k is an ancestor of l or k == l
};

void base_fk(const H1Model& cm, const double* q, BaseKin& k) {
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(k.E[0], I3, sizeof(I3));
  k.r[0][0] = k.r[0][1] = k.r[0][2] = 0;
  for (int b = 0; b < H1_NB; ++b) {
    if (b > 0) {
      int p = cm.parent[b];
      double off[3];
      matvec3(k.E[p], cm.pos[b], off);
      for (int i = 0; i < 3; ++i) k.r[b][i] = k.r[p][i] + off[i];
      matmul3(k.E[p], cm.rfix[b], k.E[b]);
      rot_axis_right(k.E[b], cm.axis[b], std::sin(q[6 + b]), std::cos(q[6 + b]));
      for (int i = 0; i < 3; ++i) k.a[b][i] = k.E[b][3 * i + cm.axis[b]];
    }
    double lc[3];
    matvec3(k.E[b], cm.ipos[b], lc);
    for (int i = 0; i < 3; ++i) k.c[b][i] = k.r[b][i] + lc[i];
  }
  for (int l = 0; l < H1_NB; ++l) {
    for (int kk = 0; kk < H1_NB; ++kk) k.anc[kk][l] = false;
    for (int a = l; a >= 0; a = cm.parent[a]) k.anc[a][l] = true;
  }
}

struct PointSet {  // joint index j = body index (1..19); index 0 unused
  double rr[3], rj[H1_NB][3], rjk[H1_NB][H1_NB][3];
  double u[3], uth[H1_NB][3], uthth[H1_NB][H1_NB][3];
};

void point_set(const H1Model& cm, const BaseKin& k, const double* wt, const double (*pt)[3], const double* vt,
               PointSet& ps) {
  double Hs[H1_NB][3], Ws[H1_NB];
  for (int b = 0; b < H1_NB; ++b) { Ws[b] = wt[b]; for (int i = 0; i < 3; ++i) Hs[b][i] = wt[b] * pt[b][i]; }
  for (int b = H1_NB - 1; b >= 1; --b) {
    int p = cm.parent[b];
    Ws[p] += Ws[b];
    for (int i = 0; i < 3; ++i) Hs[p][i] += Hs[b][i];
  }
  for (int i = 0; i < 3; ++i) ps.rr[i] = Hs[0][i];
  std::memset(ps.rj, 0, sizeof(ps.rj)); std::memset(ps.rjk, 0, sizeof(ps.rjk));
  std::memset(ps.uth, 0, sizeof(ps.uth)); std::memset(ps.uthth, 0, sizeof(ps.uthth));
  for (int l = 1; l < H1_NB; ++l) {
    double mu[3];
    for (int i = 0; i < 3; ++i) mu[i] = Hs[l][i] - Ws[l] * k.r[l][i];
    cross3(k.a[l], mu, ps.rj[l]);
  }
  for (int l = 1; l < H1_NB; ++l)
    for (int kk = 1; kk < H1_NB; ++kk)
      if (k.anc[kk][l]) {
        cross3(k.a[kk], ps.rj[l], ps.rjk[kk][l]);
        for (int i = 0; i < 3; ++i) ps.rjk[l][kk][i] = ps.rjk[kk][l][i];
      }
  if (!vt) return;
  const double* vb = vt; const double* wb = vt + 3;
  double Om[H1_NB][3], Dsub[H1_NB][3], D[H1_NB][3];
  Om[0][0] = Om[0][1] = Om[0][2] = 0;
  for (int b = 1; b < H1_NB; ++b)
    for (int i = 0; i < 3; ++i) Om[b][i] = Om[cm.parent[b]][i] + vt[5 + b] * k.a[b][i];
  std::memset(Dsub, 0, sizeof(Dsub));
  for (int b = H1_NB - 1; b >= 1; --b) {
    for (int i = 0; i < 3; ++i) { D[b][i] = Dsub[b][i]; Dsub[b][i] += vt[5 + b] * ps.rj[b][i]; }
    for (int i = 0; i < 3; ++i) Dsub[cm.parent[b]][i] += Dsub[b][i];
  }
  double wr[3];
  cross3(wb, ps.rr, wr);
  for (int i = 0; i < 3; ++i) ps.u[i] = vb[i] + wr[i] + Dsub[0][i];
  for (int kk = 1; kk < H1_NB; ++kk) {
    double wk[3] = {wb[0] + Om[kk][0], wb[1] + Om[kk][1], wb[2] + Om[kk][2]}, t1[3], t2[3];
    cross3(wk, ps.rj[kk], t1);
    cross3(k.a[kk], D[kk], t2);
    for (int i = 0; i < 3; ++i) ps.uth[kk][i] = t1[i] + t2[i];
    for (int l = 1; l < H1_NB; ++l) {
      if (!k.anc[kk][l]) continue;
      double dO[3] = {Om[l][0] - Om[kk][0], Om[l][1] - Om[kk][1], Om[l][2] - Om[kk][2]};
      double s1[3], s2[3], s3[3], s4[3], s5[3];
      cross3(wk, ps.rjk[kk][l], s1);
      cross3(dO, ps.rj[l], s2); cross3(k.a[kk], s2, s3);
      cross3(k.a[l], D[l], s4); cross3(k.a[kk], s4, s5);
      for (int i = 0; i < 3; ++i) ps.uthth[kk][l][i] = ps.uthth[l][kk][i] = s1[i] + s3[i] + s5[i];
    }
  }
}

void dR_dxi(const double* xi, int a, double* D) {  // derivative of the quaternion polynomial rotation
  double x = xi[0], y = xi[1], z = xi[2], w = xi[3];
  switch (a) {
    case 0: { double t[9] = {0, 2 * y, 2 * z, 2 * y, -4 * x, -2 * w, 2 * z, 2 * w, -4 * x}; std::memcpy(D, t, sizeof(t)); break; }
    case 1: { double t[9] = {-4 * y, 2 * x, 2 * w, 2 * x, 0, 2 * z, -2 * w, 2 * z, -4 * y}; std::memcpy(D, t, sizeof(t)); break; }
    case 2: { double t[9] = {-4 * z, -2 * w, 2 * x, 2 * w, -4 * z, 2 * y, 2 * x, 2 * y, 0}; std::memcpy(D, t, sizeof(t)); break; }
    default: { double t[9] = {0, -2 * z, 2 * y, 2 * z, 0, -2 * x, -2 * y, 2 * x, 0}; std::memcpy(D, t, sizeof(t)); break; }
  }
}

// Vector function with Jacobian (3x51, row-major J[c][i]) and Hessian contraction sum_c lam_c d2F_c.
struct VecFun {
  double val[3];
  double J[3][H1_NX];
};

struct Ctx {
  const double* xt;  // Pinocchio-ordered state
  double R[9], Ra[4][9], Rab[4][4][9];
};

void make_ctx(const double* xt, Ctx& c) {
  c.xt = xt;
  quat_poly_rot(xt + 3, c.R);
  for (int a = 0; a < 4; ++a) dR_dxi(xt + 3, a, c.Ra[a]);
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double e[4] = {0, 0, 0, 0};
      e[b] = 1.0;
      dR_dxi(e, a, c.Rab[a][b]);  // Ra is linear in xi
    }
}

void eval_P(const Ctx& c, const PointSet& ps, VecFun& f) {
  std::memset(f.J, 0, sizeof(f.J));
  double Rr[3];
  matvec3(c.R, ps.rr, Rr);
  for (int i = 0; i < 3; ++i) { f.val[i] = c.xt[i] + Rr[i]; f.J[i][i] = 1.0; }
  for (int a = 0; a < 4; ++a) { double t[3]; matvec3(c.Ra[a], ps.rr, t); for (int i = 0; i < 3; ++i) f.J[i][3 + a] = t[i]; }
  for (int j = 1; j < H1_NB; ++j) { double t[3]; matvec3(c.R, ps.rj[j], t); for (int i = 0; i < 3; ++i) f.J[i][6 + j] = t[i]; }
}
void hess_P(const Ctx& c, const PointSet& ps, const double* lam, double* H) {  // H += sum_c lam_c d2P_c (col-major 51x51)
  auto add = [&](int i, int j, double v) { H[j * H1_NX + i] += v; if (i != j) H[i * H1_NX + j] += v; };
  double Rtl[3];
  matTvec3(c.R, lam, Rtl);
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b <= a; ++b) { double t[3]; matvec3(c.Rab[a][b], ps.rr, t); add(3 + a, 3 + b, dot3(lam, t)); }
  for (int a = 0; a < 4; ++a)
    for (int j = 1; j < H1_NB; ++j) { double t[3]; matvec3(c.Ra[a], ps.rj[j], t); add(3 + a, 6 + j, dot3(lam, t)); }
  for (int j = 1; j < H1_NB; ++j)
    for (int k = 1; k <= j; ++k) add(6 + j, 6 + k, dot3(Rtl, ps.rjk[j][k]));
}
void eval_U(const Ctx& c, const PointSet& ps, VecFun& f) {
  std::memset(f.J, 0, sizeof(f.J));
  matvec3(c.R, ps.u, f.val);
  for (int a = 0; a < 4; ++a) { double t[3]; matvec3(c.Ra[a], ps.u, t); for (int i = 0; i < 3; ++i) f.J[i][3 + a] = t[i]; }
  for (int j = 1; j < H1_NB; ++j) { double t[3]; matvec3(c.R, ps.uth[j], t); for (int i = 0; i < 3; ++i) f.J[i][6 + j] = t[i]; }
  for (int m = 0; m < 3; ++m) {
    double e[3] = {0, 0, 0}, er[3], t[3];
    e[m] = 1.0;
    for (int i = 0; i < 3; ++i) f.J[i][H1_NQ + m] = c.R[3 * i + m];
    cross3(e, ps.rr, er);
    matvec3(c.R, er, t);
    for (int i = 0; i < 3; ++i) f.J[i][H1_NQ + 3 + m] = t[i];
  }
  for (int j = 1; j < H1_NB; ++j) { double t[3]; matvec3(c.R, ps.rj[j], t); for (int i = 0; i < 3; ++i) f.J[i][H1_NQ + 5 + j] = t[i]; }
}
void hess_U(const Ctx& c, const PointSet& ps, const double* lam, double* H) {
  auto add = [&](int i, int j, double v) { H[j * H1_NX + i] += v; if (i != j) H[i * H1_NX + j] += v; };
  double Rtl[3];
  matTvec3(c.R, lam, Rtl);
  for (int a = 0; a < 4; ++a) {
    double Ral[3];  // Ra^T lam
    matTvec3(c.Ra[a], lam, Ral);
    for (int b = 0; b <= a; ++b) { double t[3]; matvec3(c.Rab[a][b], ps.u, t); add(3 + a, 3 + b, dot3(lam, t)); }
    for (int k = 1; k < H1_NB; ++k) add(3 + a, 6 + k, dot3(Ral, ps.uth[k]));
    for (int m = 0; m < 3; ++m) {
      double e[3] = {0, 0, 0}, er[3];
      e[m] = 1.0;
      add(3 + a, H1_NQ + m, Ral[m]);
      cross3(e, ps.rr, er);
      add(3 + a, H1_NQ + 3 + m, dot3(Ral, er));
    }
    for (int j = 1; j < H1_NB; ++j) add(3 + a, H1_NQ + 5 + j, dot3(Ral, ps.rj[j]));
  }
  for (int k = 1; k < H1_NB; ++k) {
    for (int l = 1; l <= k; ++l) add(6 + k, 6 + l, dot3(Rtl, ps.uthth[k][l]));
    for (int m = 0; m < 3; ++m) {
      double e[3] = {0, 0, 0}, er[3];
      e[m] = 1.0;
      cross3(e, ps.rj[k], er);
      add(6 + k, H1_NQ + 3 + m, dot3(Rtl, er));
    }
    for (int j = 1; j < H1_NB; ++j) add(6 + k, H1_NQ + 5 + j, dot3(Rtl, ps.rjk[j][k]));
  }
}

void add_JtJ(const double* Ja, const double* Jb, double s, double* H) {  // H += s * Ja^T Jb (rows of J)
  for (int j = 0; j < H1_NX; ++j) {
    if (Jb[j] == 0.0) continue;
    for (int i = 0; i < H1_NX; ++i) H[j * H1_NX + i] += s * Ja[i] * Jb[j];
  }
}

}  // namespace

void cost_term_analytic(const H1Model& cm, int term, int ee, const double* x_mj, const double* target, double w,
                        double* g, double* H) {
  double xt[H1_NX];
  to_pin(x_mj, xt);
  if (term == TERM_UPRIGHT) {
    const double* s = xt + 3;  // roles (qw,qx,qy,qz) = s0..s3 (Q4)
    double z[3] = {2 * (s[1] * s[3] + s[0] * s[2]), 2 * (s[2] * s[3] - s[0] * s[1]), 1 - 2 * (s[1] * s[1] + s[2] * s[2])};
    double r[3] = {z[0], z[1], z[2] - 1.0};
    double J[3][4] = {{2 * s[2], 2 * s[3], 2 * s[0], 2 * s[1]}, {-2 * s[1], -2 * s[0], 2 * s[3], 2 * s[2]}, {0, -4 * s[1], -4 * s[2], 0}};
    double Hz[3][4][4];
    std::memset(Hz, 0, sizeof(Hz));
    Hz[0][0][2] = Hz[0][2][0] = 2; Hz[0][1][3] = Hz[0][3][1] = 2;
    Hz[1][0][1] = Hz[1][1][0] = -2; Hz[1][2][3] = Hz[1][3][2] = 2;
    Hz[2][1][1] = -4; Hz[2][2][2] = -4;
    for (int a = 0; a < 4; ++a) {
      g[3 + a] += w * (J[0][a] * r[0] + J[1][a] * r[1] + J[2][a] * r[2]);
      for (int b = 0; b < 4; ++b) {
        double v = 0;
        for (int c = 0; c < 3; ++c) v += J[c][a] * J[c][b] + r[c] * Hz[c][a][b];
        H[(3 + b) * H1_NX + 3 + a] += w * v;
      }
    }
    return;
  }
  BaseKin bk;
  base_fk(cm, xt, bk);
  Ctx ctx;
  make_ctx(xt, ctx);
  double wt[H1_NB];
  double pt[H1_NB][3];
  bool is_frame = (term == TERM_EE_POS || term == TERM_EE_VEL);
  for (int b = 0; b < H1_NB; ++b) {
    if (is_frame) { wt[b] = (b == cm.foot_body[ee]) ? 1.0 : 0.0; for (int i = 0; i < 3; ++i) pt[b][i] = bk.r[b][i]; }
    else { wt[b] = cm.mass[b] / cm.total_mass; for (int i = 0; i < 3; ++i) pt[b][i] = bk.c[b][i]; }
  }
  bool need_v = (term == TERM_COM_VEL || term == TERM_EE_VEL || term == TERM_BALANCE);
  static thread_local PointSet ps;
  point_set(cm, bk, wt, pt, need_v ? xt + H1_NQ : nullptr, ps);
  VecFun P, U;
  if (term == TERM_COM || term == TERM_EE_POS || term == TERM_COM_VEL || term == TERM_EE_VEL) {
    bool pos = (term == TERM_COM || term == TERM_EE_POS);
    VecFun& F = pos ? P : U;
    if (pos) eval_P(ctx, ps, F); else eval_U(ctx, ps, F);
    double lam[3];
    for (int c = 0; c < 3; ++c) lam[c] = 2.0 * w * (F.val[c] - target[c]);
    for (int i = 0; i < H1_NX; ++i) g[i] += lam[0] * F.J[0][i] + lam[1] * F.J[1][i] + lam[2] * F.J[2][i];
    for (int c = 0; c < 3; ++c) add_JtJ(F.J[c], F.J[c], 2.0 * w, H);
    if (pos) hess_P(ctx, ps, lam, H); else hess_U(ctx, ps, lam, H);
    return;
  }
  // balance
  eval_P(ctx, ps, P);
  eval_U(ctx, ps, U);
  const double g9 = 9.81;
  double sg = std::sqrt(P.val[2] / g9), sg1 = 1.0 / (2.0 * g9 * sg), sg2 = -1.0 / (4.0 * g9 * g9 * sg * sg * sg);
  double rho[2], Jr[2][H1_NX];
  for (int k = 0; k < 2; ++k) {
    rho[k] = P.val[k] + sg * U.val[k] - target[k];
    for (int i = 0; i < H1_NX; ++i) Jr[k][i] = P.J[k][i] + sg * U.J[k][i] + U.val[k] * sg1 * P.J[2][i];
  }
  std::vector<double> Hl(H1_NX * H1_NX, 0.0);
  for (int i = 0; i < H1_NX; ++i) g[i] += w * (rho[0] * Jr[0][i] + rho[1] * Jr[1][i]);
  for (int k = 0; k < 2; ++k) {
    add_JtJ(Jr[k], Jr[k], w, Hl.data());
    add_JtJ(U.J[k], P.J[2], w * rho[k] * sg1, Hl.data());
    add_JtJ(P.J[2], U.J[k], w * rho[k] * sg1, Hl.data());
    add_JtJ(P.J[2], P.J[2], w * rho[k] * U.val[k] * sg2, Hl.data());
  }
  double lamP[3] = {w * rho[0], w * rho[1], w * sg1 * (rho[0] * U.val[0] + rho[1] * U.val[1])};
  double lamU[3] = {w * sg * rho[0], w * sg * rho[1], 0.0};
  hess_P(ctx, ps, lamP, Hl.data());
  hess_U(ctx, ps, lamU, Hl.data());
  for (int j = 0; j < H1_NX; ++j)  // 0.5 (H + H^T), derivatives.cpp:796
    for (int i = 0; i < H1_NX; ++i) H[j * H1_NX + i] += 0.5 * (Hl[j * H1_NX + i] + Hl[i * H1_NX + j]);
}

// ---------------------------------------------------------------------------------------------
// Part 3: soft limit penalties, robot_utils.cpp:615-778 (10% margin on each side, w * violation^2)
// ---------------------------------------------------------------------------------------------
static inline void limit_1d(double val, double lo, double hi, double w, double* c, double* g, double* hdiag) {
  double margin = 0.1 * (hi - lo), lo_s = lo + margin, hi_s = hi - margin;
  if (val > hi_s) { double viol = val - hi_s; *c += w * viol * viol; *g += 2.0 * w * viol; }
  if (val < lo_s) { double viol = lo_s - val; *c += w * viol * viol; *g += -2.0 * w * viol; }
  if (val > hi_s || val < lo_s) *hdiag += 2.0 * w;
}

double limit_cost(const H1Model& md, const H1Weights& w, const double* x, const double* u) {
  double c = 0, g = 0, h = 0;
  for (int i = 0; i < H1_NU; ++i) limit_1d(u[i], md.ctrl_range[i][0], md.ctrl_range[i][1], w.w_control_limits, &c, &g, &h);
  for (int i = 0; i < H1_NU; ++i) {
    double lo = md.jnt_range[i][0], hi = md.jnt_range[i][1];
    if (std::isfinite(lo) && std::isfinite(hi) && lo < hi) limit_1d(x[7 + i], lo, hi, w.w_joint_limits, &c, &g, &h);
  }
  return c;
}

void limit_derivs(const H1Model& md, const H1Weights& w, const double* x, const double* u, double* gx, double* gu,
                  double* hxx_diag, double* huu_diag) {
  double c = 0;
  for (int i = 0; i < H1_NX; ++i) { gx[i] = 0; hxx_diag[i] = 0; }
  for (int i = 0; i < H1_NU; ++i) { gu[i] = 0; huu_diag[i] = 0; }
  for (int i = 0; i < H1_NU; ++i) limit_1d(u[i], md.ctrl_range[i][0], md.ctrl_range[i][1], w.w_control_limits, &c, &gu[i], &huu_diag[i]);
  for (int i = 0; i < H1_NU; ++i) {
    double lo = md.jnt_range[i][0], hi = md.jnt_range[i][1];
    if (std::isfinite(lo) && std::isfinite(hi) && lo < hi) limit_1d(x[7 + i], lo, hi, w.w_joint_limits, &c, &gx[7 + i], &hxx_diag[7 + i]);
  }
}

}  // namespace orc
