// Riccati backward pass, one CTA per MPC instance, sequential in the knot index, all operands of a knot in
// shared memory (reference: iLQR::backwardPass, /root/reference/src/ilqr/ilqr.cpp:250-309):
//   Qx = lx + A'Vx, Qu = lu + B'Vx, Qxx = lxx + A'VxxA, Quu = luu + B'VxxB + lambda I, Qxu = A'VxxB (lxu == 0),
//   LLT test on Quu (+1e-4 I once on failure, quirk Q9), K = -Quu^-1 Qxu', k = -Quu^-1 Qu (LDLT with
//   largest-|diagonal| pivoting like Eigen::LDLT), Vx = Qx + K'Quu k + K'Qu + Qxu k,
//   Vxx = sym(Qxx + K'QuuK + K'Qxu' + Qxu K).
// The two large products share W = Vxx [A|B] (51x70): [Qxx|Qxu] = A' W, Quu = B' W[:,51:].
// fp64 FMA register-tiled GEMMs on shared-memory operands (4x4 micro-tiles).
#pragma once
#include "h1_common.cuh"

namespace h1 {

constexpr int RIC_THREADS = 256;
constexpr int NXU = NX + NU;  // 70

// C(m x n) = op(A) * B with op(A) = A^T if TA (A stored k x m) else A (m x k); column-major, shared memory.
template <bool TA>
__device__ __forceinline__ void gemm_smem(int m, int n, int k, const double* __restrict__ A, int lda,
                                          const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc) {
  const int tm = (m + 3) >> 2, tn = (n + 3) >> 2;
  for (int tile = threadIdx.x; tile < tm * tn; tile += blockDim.x) {
    const int ti = tile % tm, tj = tile / tm;
    const int i0 = ti * 4, j0 = tj * 4;
    int ri[4], cj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { ri[r] = min(i0 + r, m - 1); cj[r] = min(j0 + r, n - 1); }
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int kk = 0; kk < k; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = TA ? A[ri[r] * lda + kk] : A[kk * lda + ri[r]];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = B[cj[c] * ldb + kk];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (i0 + r < m && j0 + c < n) C[(j0 + c) * ldc + i0 + r] = acc[r][c];
  }
}

struct RiccatiSmem {
  double Vxx[NX * NX];
  double AB[NX * NXU];     // [A | B]
  double W[NX * NXU];      // Vxx [A | B]; later reused as scratch (QuuK, T)
  double Qx_[NX * NXU];    // [Qxx | Qxu]
  double Quu[NU * NU];
  double Lf[NU * NU];      // permuted LDLT factor (unit lower), D on the diagonal slot array below
  double Kt[NU * (NX + 1)];// solves: columns 0..50 -> K(:,j), column 51 -> kff
  double Vx[NX], Qx[NX], Qu[NU], D[NU], tmp[NX];
  int perm[NU];
  int llt_fail;
};

// Cholesky positive-definiteness test of Quu (Eigen::LLT info()); one warp, lane <-> row. Scratch in Lf.
__device__ __forceinline__ void quu_llt_check(RiccatiSmem& s) {
  const int lane = threadIdx.x;  // called by warp 0 only
  const int n = NU;
  for (int e = lane; e < n * n; e += 32) s.Lf[e] = s.Quu[e];
  if (lane == 0) s.llt_fail = 0;
  __syncwarp();
  for (int j = 0; j < n; ++j) {
    const double d = s.Lf[j * n + j];
    if (!(d > 0.0)) { if (lane == 0) s.llt_fail = 1; break; }
    const double sd = sqrt(d);
    double lij = 0.0;
    if (lane > j && lane < n) { lij = s.Lf[j * n + lane] / sd; s.Lf[j * n + lane] = lij; }
    __syncwarp();
    // trailing update of row `lane`: M[lane][c] -= l[lane][j] * l[c][j], c in (j, lane]
    if (lane > j && lane < n)
      for (int c = j + 1; c <= lane; ++c) s.Lf[c * n + lane] -= lij * s.Lf[j * n + c];
    __syncwarp();
  }
  __syncwarp();
}

// LDL^T with symmetric pivoting by largest |original diagonal| (selection order as Eigen::LDLT), one warp.
__device__ __forceinline__ void quu_ldlt(RiccatiSmem& s) {
  const int lane = threadIdx.x;
  const int n = NU;
  if (lane == 0) {
    int p[NU];
    for (int i = 0; i < n; ++i) p[i] = i;
    for (int k = 0; k < n; ++k) {
      int best = k;
      double bv = fabs(s.Quu[p[k] * n + p[k]]);
      for (int i = k + 1; i < n; ++i) { const double v = fabs(s.Quu[p[i] * n + p[i]]); if (v > bv) { bv = v; best = i; } }
      const int t = p[k]; p[k] = p[best]; p[best] = t;
    }
    for (int i = 0; i < n; ++i) s.perm[i] = p[i];
  }
  __syncwarp();
  for (int e = lane; e < n * n; e += 32) { const int i = e % n, j = e / n; s.Lf[e] = s.Quu[s.perm[j] * n + s.perm[i]]; }
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    if (lane == k) {
      double d = s.Lf[k * n + k];
      for (int c = 0; c < k; ++c) d -= s.Lf[c * n + k] * s.Lf[c * n + k] * s.D[c];
      s.D[k] = d;
    }
    __syncwarp();
    if (lane > k && lane < n) {
      double acc = s.Lf[k * n + lane];
      for (int c = 0; c < k; ++c) acc -= s.Lf[c * n + lane] * s.Lf[c * n + k] * s.D[c];
      const double d = s.D[k];
      s.Lf[k * n + lane] = (fabs(d) > 0.0) ? acc / d : 0.0;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(RIC_THREADS)
k_backward(int N, const int* __restrict__ mask, const double* __restrict__ lambda, const double* __restrict__ A,
           const double* __restrict__ Bm, const double* __restrict__ lx, const double* __restrict__ lu,
           const double* __restrict__ lxx, const double* __restrict__ luu, double* __restrict__ K,
           double* __restrict__ kff, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RiccatiSmem& s = *reinterpret_cast<RiccatiSmem*>(smem_raw);
  const int inst = blockIdx.x;
  if (mask && !mask[inst]) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double lam = lambda[inst];
  const double* lxN = lx + ((size_t)inst * (N + 1) + N) * NX;
  const double* lxxN = lxx + ((size_t)inst * (N + 1) + N) * NX * NX;
  for (int i = tid; i < NX; i += nt) s.Vx[i] = lxN[i];
  for (int i = tid; i < NX * NX; i += nt) s.Vxx[i] = lxxN[i];
  bool nonfinite = false;
  for (int t = N - 1; t >= 0; --t) {
    const double* At = A + ((size_t)inst * N + t) * NX * NX;
    const double* Bt = Bm + ((size_t)inst * N + t) * NX * NU;
    const double* lxt = lx + ((size_t)inst * (N + 1) + t) * NX;
    const double* lut = lu + ((size_t)inst * N + t) * NU;
    const double* lxxt = lxx + ((size_t)inst * (N + 1) + t) * NX * NX;
    const double* luut = luu + ((size_t)inst * N + t) * NU * NU;
    for (int i = tid; i < NX * NX; i += nt) s.AB[i] = At[i];
    for (int i = tid; i < NX * NU; i += nt) s.AB[NX * NX + i] = Bt[i];
    __syncthreads();
    // W = Vxx [A|B]
    gemm_smem<false>(NX, NXU, NX, s.Vxx, NX, s.AB, NX, s.W, NX);
    // Qx = lx + A'Vx, Qu = lu + B'Vx
    for (int i = tid; i < NXU; i += nt) {
      double acc = 0.0;
      const double* col = s.AB + i * NX;
      for (int l = 0; l < NX; ++l) acc += col[l] * s.Vx[l];
      if (i < NX) s.Qx[i] = lxt[i] + acc; else s.Qu[i - NX] = lut[i - NX] + acc;
    }
    __syncthreads();
    // [Qxx | Qxu] = A' W ; Quu = B' W[:, 51:]
    gemm_smem<true>(NX, NXU, NX, s.AB, NX, s.W, NX, s.Qx_, NX);
    gemm_smem<true>(NU, NU, NX, s.AB + NX * NX, NX, s.W + NX * NX, NX, s.Quu, NU);
    __syncthreads();
    for (int i = tid; i < NX * NX; i += nt) s.Qx_[i] += lxxt[i];
    for (int i = tid; i < NU * NU; i += nt) s.Quu[i] += luut[i] + ((i % NU == i / NU) ? lam : 0.0);
    __syncthreads();
    if (tid < 32) quu_llt_check(s);
    __syncthreads();
    if (s.llt_fail) {
      for (int i = tid; i < NU; i += nt) s.Quu[i * NU + i] += 1e-4;
      __syncthreads();
    }
    if (tid < 32) quu_ldlt(s);
    __syncthreads();
    // solves: rhs r < 51 -> column r of Qxu' (= row r of Qxu), rhs 51 -> Qu ; result negated
    if (tid <= NX) {
      const int r = tid;
      double y[NU];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const int pi = s.perm[i];
        y[i] = (r < NX) ? s.Qx_[(NX + pi) * NX + r] : s.Qu[pi];
      }
#pragma unroll
      for (int i = 0; i < NU; ++i)
#pragma unroll
        for (int c = 0; c < i; ++c) y[i] -= s.Lf[c * NU + i] * y[c];
#pragma unroll
      for (int i = 0; i < NU; ++i) y[i] = (fabs(s.D[i]) > 2.2250738585072014e-308) ? y[i] / s.D[i] : 0.0;
#pragma unroll
      for (int i = NU - 1; i >= 0; --i)
#pragma unroll
        for (int c = i + 1; c < NU; ++c) y[i] -= s.Lf[i * NU + c] * y[c];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const double v = -y[i];
        if (!isfinite(v)) nonfinite = true;
        s.Kt[r * NU + s.perm[i]] = v;
      }
    }
    __syncthreads();
    // write gains (K column-major 19x51 == Kt columns 0..50; kff = column 51)
    double* Kt_g = K + ((size_t)inst * N + t) * NU * NX;
    double* kf_g = kff + ((size_t)inst * N + t) * NU;
    for (int i = tid; i < NU * NX; i += nt) Kt_g[i] = s.Kt[i];
    for (int i = tid; i < NU; i += nt) kf_g[i] = s.Kt[NX * NU + i];
    // QuuK (19x51) into W scratch, Quu k into tmp
    double* QuuK = s.W;
    double* T3 = s.W + NU * NX;  // Qxu K (51x51)
    gemm_smem<false>(NU, NX, NU, s.Quu, NU, s.Kt, NU, QuuK, NU);
    gemm_smem<false>(NX, NX, NU, s.Qx_ + NX * NX, NX, s.Kt, NU, T3, NX);
    if (tid < NU) {
      double acc = 0.0;
      for (int l = 0; l < NU; ++l) acc += s.Quu[l * NU + tid] * s.Kt[NX * NU + l];
      s.tmp[tid] = acc;
    }
    __syncthreads();
    // Vx = Qx + K'(Quu k) + K'Qu + Qxu k
    if (tid < NX) {
      double a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int l = 0; l < NU; ++l) {
        const double kli = s.Kt[tid * NU + l];
        a1 += kli * s.tmp[l]; a2 += kli * s.Qu[l]; a3 += s.Qx_[(NX + l) * NX + tid] * s.Kt[NX * NU + l];
      }
      s.Vx[tid] = s.Qx[tid] + a1 + a2 + a3;
    }
    // T(i,j) = Qxx + K'QuuK + (QxuK)' + QxuK, written over the Qxx block
    for (int e = tid; e < NX * NX; e += nt) {
      const int i = e % NX, j = e / NX;
      double acc = 0.0;
      for (int l = 0; l < NU; ++l) acc += s.Kt[i * NU + l] * QuuK[j * NU + l];
      s.Qx_[e] = s.Qx_[e] + acc + T3[i * NX + j] + T3[e];
    }
    __syncthreads();
    for (int e = tid; e < NX * NX; e += nt) {
      const int i = e % NX, j = e / NX;
      s.Vxx[e] = 0.5 * (s.Qx_[e] + s.Qx_[i * NX + j]);
    }
    __syncthreads();
  }
  if (nonfinite && status) status[inst] = 1;
}

}  // namespace h1
