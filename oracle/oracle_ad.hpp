// TEST INFRASTRUCTURE ONLY — second-order forward-mode AD scalar (value, gradient, packed Hessian).
// Plays the role CasADi SX::gradient / SX::jacobian play in the reference
// (/root/reference/src/common/derivatives.cpp:95-106,147-157,198-199,725-726): exact first and
// second derivatives of a scalar expression w.r.t. the 51 raw state entries, all terms kept.
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

template <int NV> struct D2 {
  static constexpr int NH = NV * (NV + 1) / 2;
  double v;
  bool c;  // constant: g and h are identically zero (and may hold garbage)
  double g[NV];
  double h[NH];  // lower triangle, idx(i,j) = i(i+1)/2 + j, i >= j
  D2() : v(0), c(true) {}
  D2(double x) : v(x), c(true) {}
  static D2 var(double x, int i) {
    D2 r;
    r.v = x; r.c = false;
    std::memset(r.g, 0, sizeof(r.g)); std::memset(r.h, 0, sizeof(r.h));
    r.g[i] = 1.0;
    return r;
  }
  double H(int i, int j) const { if (c) return 0; return i >= j ? h[i * (i + 1) / 2 + j] : h[j * (j + 1) / 2 + i]; }
  double G(int i) const { return c ? 0.0 : g[i]; }
};

template <int NV> inline D2<NV> unary(const D2<NV>& a, double f, double f1, double f2) {
  D2<NV> r;
  r.v = f;
  if (a.c) { r.c = true; return r; }
  r.c = false;
  for (int i = 0; i < NV; ++i) r.g[i] = f1 * a.g[i];
  int k = 0;
  for (int i = 0; i < NV; ++i) {
    double gi = f2 * a.g[i];
    for (int j = 0; j <= i; ++j, ++k) r.h[k] = f1 * a.h[k] + gi * a.g[j];
  }
  return r;
}

template <int NV> inline D2<NV> operator+(const D2<NV>& a, const D2<NV>& b) {
  if (a.c && b.c) return D2<NV>(a.v + b.v);
  if (a.c) { D2<NV> r = b; r.v += a.v; return r; }
  if (b.c) { D2<NV> r = a; r.v += b.v; return r; }
  D2<NV> r; r.c = false; r.v = a.v + b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < D2<NV>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
  return r;
}
template <int NV> inline D2<NV> operator-(const D2<NV>& a) {
  if (a.c) return D2<NV>(-a.v);
  D2<NV> r; r.c = false; r.v = -a.v;
  for (int i = 0; i < NV; ++i) r.g[i] = -a.g[i];
  for (int i = 0; i < D2<NV>::NH; ++i) r.h[i] = -a.h[i];
  return r;
}
template <int NV> inline D2<NV> operator-(const D2<NV>& a, const D2<NV>& b) { return a + (-b); }
template <int NV> inline D2<NV> operator*(const D2<NV>& a, const D2<NV>& b) {
  if (a.c && b.c) return D2<NV>(a.v * b.v);
  if (a.c) return unary(b, a.v * b.v, a.v, 0.0);
  if (b.c) return unary(a, a.v * b.v, b.v, 0.0);
  D2<NV> r; r.c = false; r.v = a.v * b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
  int k = 0;
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j <= i; ++j, ++k)
      r.h[k] = a.v * b.h[k] + b.v * a.h[k] + a.g[i] * b.g[j] + b.g[i] * a.g[j];
  return r;
}
template <int NV> inline D2<NV> inv(const D2<NV>& a) { double i = 1.0 / a.v; return unary(a, i, -i * i, 2 * i * i * i); }
template <int NV> inline D2<NV> operator/(const D2<NV>& a, const D2<NV>& b) { return a * inv(b); }
template <int NV> inline D2<NV> sqrt(const D2<NV>& a) { double s = std::sqrt(a.v); return unary(a, s, 0.5 / s, -0.25 / (s * a.v)); }
template <int NV> inline D2<NV> sin(const D2<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return unary(a, s, c, -s); }
template <int NV> inline D2<NV> cos(const D2<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return unary(a, c, -s, -c); }
// mixed with double
template <int NV> inline D2<NV> operator+(const D2<NV>& a, double b) { return a + D2<NV>(b); }
template <int NV> inline D2<NV> operator+(double b, const D2<NV>& a) { return a + D2<NV>(b); }
template <int NV> inline D2<NV> operator-(const D2<NV>& a, double b) { return a + D2<NV>(-b); }
template <int NV> inline D2<NV> operator-(double b, const D2<NV>& a) { return D2<NV>(b) - a; }
template <int NV> inline D2<NV> operator*(const D2<NV>& a, double b) { return a * D2<NV>(b); }
template <int NV> inline D2<NV> operator*(double b, const D2<NV>& a) { return a * D2<NV>(b); }
template <int NV> inline D2<NV> operator/(const D2<NV>& a, double b) { return a * D2<NV>(1.0 / b); }

// ---- first-order forward-mode dual with N tangent directions (exact Jacobian of f_D in one pass) ----
template <int N> struct D1 {
  double v;
  double d[N];
  D1() : v(0) { for (int i = 0; i < N; ++i) d[i] = 0; }
  D1(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0; }
  static D1 var(double x, int i) { D1 r(x); r.d[i] = 1.0; return r; }
};
template <int N> inline D1<N> operator+(const D1<N>& a, const D1<N>& b) { D1<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline D1<N> operator-(const D1<N>& a, const D1<N>& b) { D1<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline D1<N> operator-(const D1<N>& a) { D1<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline D1<N> operator*(const D1<N>& a, const D1<N>& b) { D1<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v; return r; }
template <int N> inline D1<N> operator/(const D1<N>& a, const D1<N>& b) { D1<N> r; r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int N> inline D1<N> operator+(const D1<N>& a, double b) { D1<N> r = a; r.v += b; return r; }
template <int N> inline D1<N> operator+(double b, const D1<N>& a) { D1<N> r = a; r.v += b; return r; }
template <int N> inline D1<N> operator-(const D1<N>& a, double b) { D1<N> r = a; r.v -= b; return r; }
template <int N> inline D1<N> operator-(double b, const D1<N>& a) { return D1<N>(b) - a; }
template <int N> inline D1<N> operator*(const D1<N>& a, double b) { D1<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> inline D1<N> operator*(double b, const D1<N>& a) { return a * b; }
template <int N> inline D1<N> operator/(const D1<N>& a, double b) { D1<N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
template <int N> inline D1<N>& operator+=(D1<N>& a, const D1<N>& b) { a = a + b; return a; }
template <int N> inline D1<N>& operator-=(D1<N>& a, const D1<N>& b) { a = a - b; return a; }
template <int N> inline bool operator<(const D1<N>& a, double b) { return a.v < b; }
template <int N> inline bool operator>(const D1<N>& a, double b) { return a.v > b; }
template <int N> inline D1<N> sqrt(const D1<N>& a) { D1<N> r; r.v = std::sqrt(a.v); for (int i = 0; i < N; ++i) r.d[i] = 0.5 * a.d[i] / r.v; return r; }
template <int N> inline D1<N> sin(const D1<N>& a) { D1<N> r; double c = std::cos(a.v); r.v = std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> inline D1<N> cos(const D1<N>& a) { D1<N> r; double s = std::sin(a.v); r.v = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = -s * a.d[i]; return r; }
inline double value_of(double x) { return x; }
template <int N> inline double value_of(const D1<N>& x) { return x.v; }

inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }

}  // namespace orc
