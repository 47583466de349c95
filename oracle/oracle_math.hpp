// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/README.md). Not linked into the product.
// Small fixed-size vector helpers, templated on the scalar so the same code runs on double and on
// the second-order AD scalar (oracle_ad.hpp).
#pragma once
#include <cmath>

namespace orc {

template <class T> inline void cross3(const T* a, const T* b, T* o) {
  T x = a[1] * b[2] - a[2] * b[1];
  T y = a[2] * b[0] - a[0] * b[2];
  T z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// o = M(3x3 row-major) * v
template <class T, class U> inline void matvec3(const T* M, const U* v, T* o) {
  T x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  T y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  T z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// o = M^T * v
template <class T, class U> inline void matTvec3(const T* M, const U* v, T* o) {
  T x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  T y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  T z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// C = A * B (3x3 row-major); A may be T, B may be a different scalar
template <class T, class U> inline void matmul3(const T* A, const U* B, T* C) {
  T out[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      out[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  for (int i = 0; i < 9; ++i) C[i] = out[i];
}
// R <- R * Rot(axis, angle) with sine s and cosine c, R row-major
template <class T> inline void rot_axis_right(T* R, int axis, const T& s, const T& c) {
  int p = (axis + 1) % 3, q = (axis + 2) % 3;  // columns mixed by an elementary rotation
  for (int i = 0; i < 3; ++i) {
    T cp = R[3 * i + p], cq = R[3 * i + q];
    R[3 * i + p] = c * cp + s * cq;
    R[3 * i + q] = c * cq - s * cp;
  }
}

}  // namespace orc
