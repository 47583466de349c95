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
  double AB[NX * NXU];     // [A | B] of the current knot; refilled with the next knot's by cp.async
  double W[NX * NXU];      // Vxx [A | B]; later reused as scratch (QuuK, Qxu K)
  double Qx_[NX * NXU];    // [Qxx | Qxu]
  double Lnext[NX * NX];   // lxx of the next knot (cp.async prefetch)
  double T1[NX * NX];      // K' Quu K
  double Quu[NU * NU];
  double Lf[NU * NU];      // permuted Quu, reduced in place to its Schur complements
  double Ls[NU * NU];      // unit-lower factor of the permuted LDL^T
  double Kt[NU * (NX + 1)];// solves: columns 0..50 -> K(:,j), column 51 -> kff
  double Vx[NX], Qx[NX], Qu[NU], D[NU], tmp[NX];
  int perm[NU];
  int not_pd;
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// LDL^T of Quu with symmetric pivoting by largest |diagonal| (Eigen::LDLT's selection rule), one warp,
// right-looking: after step k, row i holds the Schur complement. Sets not_pd when a pivot is <= 0, which for
// a symmetric matrix is equivalent to Eigen::LLT reporting failure (Sylvester's law of inertia).
__device__ __forceinline__ void quu_ldlt(RiccatiSmem& s) {
  const int lane = threadIdx.x;  // warp 0 only
  const int n = NU;
  if (lane < n) {  // rank of |Q_ii| in descending order, ties by index
    const double di = fabs(s.Quu[lane * n + lane]);
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = fabs(s.Quu[j * n + j]);
      rank += (dj > di) || (dj == di && j < lane);
    }
    s.perm[rank] = lane;
  }
  if (lane == 0) s.not_pd = 0;
  __syncwarp();
  for (int e = lane; e < n * n; e += 32) { const int i = e % n, j = e / n; s.Lf[e] = s.Quu[s.perm[j] * n + s.perm[i]]; }
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    const double d = s.Lf[k * n + k];
    if (lane == 0) { s.D[k] = d; if (!(d > 0.0)) s.not_pd = 1; }
    if (lane > k && lane < n) {
      const double pik = s.Lf[k * n + lane];                       // P(i,k)
      const double lik = (fabs(d) > 0.0) ? pik / d : 0.0;
      for (int c = k + 1; c <= lane; ++c) s.Lf[c * n + lane] -= lik * s.Lf[k * n + c];   // P(i,c) -= l_ik P(c,k)
      s.Ls[k * n + lane] = lik;   // separate array: column k of Lf is still being read by the other lanes
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
  auto prefetch = [&](int t) {  // A_t, B_t -> s.AB ; lxx_t -> s.Lnext  (8-byte async copies)
    const double* At = A + ((size_t)inst * N + t) * NX * NX;
    const double* Bt = Bm + ((size_t)inst * N + t) * NX * NU;
    const double* Lt = lxx + ((size_t)inst * (N + 1) + t) * NX * NX;
    for (int i = tid; i < NX * NX; i += nt) { cp_async8(&s.AB[i], At + i); cp_async8(&s.Lnext[i], Lt + i); }
    for (int i = tid; i < NX * NU; i += nt) cp_async8(&s.AB[NX * NX + i], Bt + i);
  };
  prefetch(N - 1);
  for (int i = tid; i < NX; i += nt) s.Vx[i] = lxN[i];
  for (int i = tid; i < NX * NX; i += nt) s.Vxx[i] = lxxN[i];
  bool nonfinite = false;
  for (int t = N - 1; t >= 0; --t) {
    const double* lxt = lx + ((size_t)inst * (N + 1) + t) * NX;
    const double* lut = lu + ((size_t)inst * N + t) * NU;
    const double* luut = luu + ((size_t)inst * N + t) * NU * NU;
    cp_async_commit_wait_all();
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
    for (int i = tid; i < NX * NX; i += nt) s.Qx_[i] += s.Lnext[i];
    for (int i = tid; i < NU * NU; i += nt) s.Quu[i] += luut[i] + ((i % NU == i / NU) ? lam : 0.0);
    __syncthreads();
    if (t > 0) prefetch(t - 1);   // s.AB / s.Lnext are free from here on; overlaps the factorisation and solves
    if (tid < 32) {
      quu_ldlt(s);
      if (s.not_pd) {             // Eigen::LLT failed: Quu += 1e-4 I once, no re-check (quirk Q9), refactor
        for (int i = tid; i < NU; i += 32) s.Quu[i * NU + i] += 1e-4;
        __syncwarp();
        quu_ldlt(s);
      }
    }
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
        for (int c = 0; c < i; ++c) y[i] -= s.Ls[c * NU + i] * y[c];
#pragma unroll
      for (int i = 0; i < NU; ++i) y[i] = (fabs(s.D[i]) > 2.2250738585072014e-308) ? y[i] / s.D[i] : 0.0;
#pragma unroll
      for (int i = NU - 1; i >= 0; --i)
#pragma unroll
        for (int c = i + 1; c < NU; ++c) y[i] -= s.Ls[i * NU + c] * y[c];
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
    // QuuK (19x51) and Qxu K (51x51) into W scratch, Quu k into tmp
    double* QuuK = s.W;
    double* T3 = s.W + NU * NX;
    gemm_smem<false>(NU, NX, NU, s.Quu, NU, s.Kt, NU, QuuK, NU);
    gemm_smem<false>(NX, NX, NU, s.Qx_ + NX * NX, NX, s.Kt, NU, T3, NX);
    if (tid < NU) {
      double acc = 0.0;
      for (int l = 0; l < NU; ++l) acc += s.Quu[l * NU + tid] * s.Kt[NX * NU + l];
      s.tmp[tid] = acc;
    }
    __syncthreads();
    // Vx = Qx + K'(Quu k) + K'Qu + Qxu k ;  T1 = K' (Quu K)
    if (tid < NX) {
      double a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int l = 0; l < NU; ++l) {
        const double kli = s.Kt[tid * NU + l];
        a1 += kli * s.tmp[l]; a2 += kli * s.Qu[l]; a3 += s.Qx_[(NX + l) * NX + tid] * s.Kt[NX * NU + l];
      }
      s.Vx[tid] = s.Qx[tid] + a1 + a2 + a3;
    }
    gemm_smem<true>(NX, NX, NU, s.Kt, NU, QuuK, NU, s.T1, NX);
    __syncthreads();
    // Vxx = sym(Qxx + K'QuuK + (QxuK)' + QxuK)
    for (int e = tid; e < NX * NX; e += nt) {
      const int i = e % NX, j = e / NX, et = i * NX + j;
      const double tij = s.Qx_[e] + s.T1[e] + T3[et] + T3[e];
      const double tji = s.Qx_[et] + s.T1[et] + T3[e] + T3[et];
      s.Vxx[e] = 0.5 * (tij + tji);
    }
    __syncthreads();
  }
  if (nonfinite && status) status[inst] = 1;
}

}  // namespace h1
