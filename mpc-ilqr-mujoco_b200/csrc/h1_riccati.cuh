// Riccati backward pass, one CTA per MPC instance, sequential in the knot index, all operands of a knot in
// shared memory (reference: iLQR::backwardPass, /root/reference/src/ilqr/ilqr.cpp:250-309):
//   Qx = lx + A'Vx, Qu = lu + B'Vx, Qxx = lxx + A'VxxA, Quu = luu + B'VxxB + lambda I, Qxu = A'VxxB (lxu == 0),
//   LLT test on Quu (+1e-4 I once on failure, quirk Q9), K = -Quu^-1 Qxu', k = -Quu^-1 Qu (LDLT with
//   largest-|diagonal| pivoting like Eigen::LDLT), Vx = Qx + K'Quu k + K'Qu + Qxu k,
//   Vxx = sym(Qxx + K'QuuK + K'Qxu' + Qxu K).
// Contractions (all on the fp64 tensor cores, mma.sync m8n8k4 = SASS DMMA, operands in shared memory):
//   G1  W = Vxx [A|B]                       51x70x51      (W is shared by the next two; its B block is computed
//                                                         first so that Quu and its factorisation start early)
//   G2  [Qxx|Qxu] = A' W                    51x70x51
//   G3  Quu = B' W[:,51:]                   19x19x51
//   G4  G = Quu K + 2 Qxu'                  19x51x19
//   G5  M = Qxx + K' G                      51x51x19      sym(M) == sym(Qxx + K'QuuK + K'Qxu' + QxuK)
// (K'Qxu' and QxuK are transposes of each other, so their sum inside the symmetrisation equals sym(2 K'Qxu');
// no algebraic cancellation is assumed, only the order of additions differs from the reference.)
// Shared memory per CTA is 107 KB so that TWO instances are resident per SM: while one is in its sequential
// section (pivoted LDLT, triangular solves) the other one keeps the tensor pipe busy.
// Leading dimensions are = 4 (mod 8) doubles: every m8n8k4 fragment load is bank-conflict free.
//
// STRUCT = true (the linearization came from the analytic kernels): f_D is semi-implicit Euler, p+ = p + h v+,
// theta+ = theta + h thetadot+, so 22 of the 26 position rows of [A|B] are  e_r' + h * (their velocity row), exactly.
// With R = the 29 rows that carry information (4 quaternion rows, 25 velocity rows), [A|B] = P + E R, P = the unit
// entries, E = selection + h * (position partner). Every contraction over the 51 state rows then runs over the 29
// reduced ones:  V [A|B] = V P + (V E) R,   A'[W|Vx] = P'[W|Vx] + R_A'(E'[W|Vx]),   B'[W_B|Vx] = R_B'(E'[W_B|Vx]);
// V E and E'W are formed in the fragment loads (one extra load + multiply-add), the P terms are one addition in the
// epilogues: 8 k-steps instead of 13 in G1, G2, G3 (1428 instead of 2028 DMMA per knot). The dense form stays for
// forward-difference and caller-supplied linearizations, which satisfy the relation only to their own accuracy.
#pragma once
#include "h1_common.cuh"
#include <type_traits>
#include <cstdio>

namespace h1 {

constexpr int RIC_THREADS = 256;
constexpr int NXU = NX + NU;  // 70
constexpr int LDX = 52;       // leading dimension of 51-row operands (row 51 is a zero pad: k runs to 52)
constexpr int LDU = 20;       // leading dimension of 19-row operands (row 19 is a zero pad: k runs to 20)

struct RiccatiSmem {
  double V[LDX * LDX];        // Vxx (pad row/column zero); from G2 on: Qxx, then M, then the next Vxx
  double AB[LDX * NXU];       // [A_t | pad | B_t | pad] exactly as they lie in global memory (dense columns, leading dimension NX,
                              // one pad double behind each: A_STRIDE + B_STRIDE doubles), refilled by two bulk copies after G2
  double W[LDX * NXU];        // Vxx [A | B]; after G2/G3: G (LDU x 51) followed by the prefetched lxx (51 x 51 dense)
  double Qxu[LDX * NU];
  double Kt[LDU * 56];        // solves: columns 0..50 -> K(:,j), column 51 -> kff; pad row 19 stays zero
  double Quu[LDU * LDU];      // pad row/column 19 stay zero
  double Ls[NU * NU];         // unit-lower factor of the permuted LDL^T
  double col[2][32];          // column exchange of the LDL^T warp (double buffered)
  double Li[LDU * LDU];       // its inverse (unit lower), row-major; pad row / column 19 stay zero
  double Vx[LDX], Qx[NX], Qu[NU], D[NU], Dinv[NU], tmp[NU];   // Vx[51] is a zero pad (Vx rides along as an extra column of W)
  double lq[NX + NU + NU * NU + 1];   // lx_t, lu_t, luu_t of the current knot (prefetched with [A|B])
  unsigned long long mbar[2];         // transaction barriers of the bulk copies: [0] lxx_t -> W, [1] [A|B] of the next knot -> AB
  int perm[NU];
};
constexpr int RIC_G_OFF = 0;              // G inside W
constexpr int RIC_LXX_OFF = LDU * NX;     // prefetched lxx inside W (dense, ld 51)
static_assert(RIC_LXX_OFF + LXX_STRIDE <= LDX * NXU, "lxx prefetch must fit behind G");
static_assert(A_STRIDE + B_STRIDE <= LDX * NXU, "[A|B] staging");
static_assert((LDX * LDX * 8) % 16 == 0 && (RIC_LXX_OFF * 8) % 16 == 0, "bulk-copy destinations are 16-byte aligned");

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
// Bulk asynchronous copy global -> shared (SASS UBLKCP: ONE instruction per matrix, executed by the copy engine) that
// signals a transaction barrier. Source, destination and size are multiples of 16 bytes (padded strides, h1_common.cuh).
// Measured on B200: the same bytes as 8-byte cp.async (global columns of a 51-row matrix are only 8-byte aligned) cost the
// issuing warps ~5 k cycles per knot — a fifth of the knot.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tRIC_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra RIC_DONE;\n\tbra RIC_WAIT;\n\tRIC_DONE:\n\t}\n"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
// reduced row index k (0..31) -> row of [A|B] / W / column of V that carries it: quaternion rows 3..6, the 25 velocity rows,
// then the zero pad row 51; and the position row whose tangent is h times velocity row k (pad row if there is none)
__device__ __forceinline__ int ric_row_v(int k) { return k < 4 ? 3 + k : min(22 + k, NX); }
__device__ __forceinline__ int ric_row_q(int k) { const int j = k - 4; return (j >= 0 && j < 3) ? j : ((j >= 6 && j < NV) ? j + 1 : NX); }
__device__ __forceinline__ bool ric_unit_row(int r) { return r < 3 || (r >= 7 && r < NQ); }   // rows of P: d q+_r / d q_r = 1

// One warp accumulates an 8 x (8 NT) strip: acc[j] += sum_k A(m0 + g, k) B(k, n0 + 8 j + g'), k < 4 ksteps.
// fa(row, k) / fb(k, col) return operand elements (they implement padding / clamping).
template <int NT, class FA, class FB>
__device__ __forceinline__ void mma_strip(int ksteps, int m0, int n0, FA fa, FB fb, double (&acc)[NT][2]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k = 4 * ks + t;
    const double a = fa(m0 + g, k);
#pragma unroll
    for (int j = 0; j < NT; ++j) dmma884(acc[j][0], acc[j][1], a, fb(k, n0 + 8 * j + g));
  }
}

// Zero-initialised strip + epilogue st(row, col, value) for every element of the 8 x (8 NT) strip.
template <int NT, class FA, class FB, class ST>
__device__ __forceinline__ void mma_strip_store(int ksteps, int m0, int n0, FA fa, FB fb, ST st) {
  double acc[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = 0.0;
  mma_strip<NT>(ksteps, m0, n0, fa, fb, acc);
  const int lane = threadIdx.x & 31, r = m0 + (lane >> 2), c0 = n0 + 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < NT; ++j) { st(r, c0 + 8 * j, acc[j][0]); st(r, c0 + 8 * j + 1, acc[j][1]); }
}

// One 8 x 8 tile whose k loop is split over NS independent accumulator chains (a dependent DMMA costs far more than
// its issue slot: short contractions that sit on the critical path of a knot are latency bound, not pipe bound).
template <int KSTEPS, int NS, class FA, class FB>
__device__ __forceinline__ void mma_tile_split(int m0, int n0, FA fa, FB fb, double& c0, double& c1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double acc[NS][2];
#pragma unroll
  for (int q = 0; q < NS; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const int k = 4 * ks + t;
    dmma884(acc[ks % NS][0], acc[ks % NS][1], fa(m0 + g, k), fb(k, n0 + g));
  }
  c0 = 0.0; c1 = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) { c0 += acc[q][0]; c1 += acc[q][1]; }
}

// 1 / d to within an ulp for a normal d (non-finite d gives a non-finite result, which marks the instance as diverged).
__device__ __forceinline__ double rcp_newton(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// LDL^T of Quu with symmetric pivoting by largest |diagonal| (Eigen::LDLT's selection rule), ONE warp, right-
// looking, register resident: lane i holds row i of the permuted matrix (lower triangle), column k of the
// current Schur complement is exchanged through shared memory. Writes perm, D, Dinv, Ls (unit-lower factor). Returns true when
// a pivot is <= 0, which for a symmetric matrix is equivalent to Eigen::LLT reporting failure (Sylvester).
__device__ __forceinline__ bool quu_ldlt(RiccatiSmem& s) {
  const int lane = threadIdx.x & 31;
  constexpr int n = NU;
  if (lane < n) {  // rank of |Q_ii| in descending order, ties by index. A NaN diagonal (diverged instance: the reference
                   // carries non-finite gains on with a warning, ilqr.cpp:290-293) sorts last, so perm is always a permutation
    auto key = [&](int j) { const double v = fabs(s.Quu[j * LDU + j]); return (v == v) ? v : -1.0; };
    const double di = key(lane);
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = key(j);
      rank += (dj > di) || (dj == di && j < lane);
    }
    s.perm[rank] = lane;
  }
  __syncwarp();
  const int li = min(lane, n - 1);
  const int pr = s.perm[li];
  double p[n];
#pragma unroll
  for (int c = 0; c < n; ++c) p[c] = s.Quu[s.perm[c] * LDU + pr];
  bool not_pd = false;
#pragma unroll
  for (int k = 0; k < n; ++k) {
    // column k of the current Schur complement goes through shared memory (one store, broadcast loads): lane c's p[k] is
    // P(c,k), lane k's is the pivot. Entries p[c] with c > lane are never read again, so they are updated unguarded.
    double* colk = s.col[k & 1];
    const double pik = p[k];                                           // P(i,k) of this lane's row
    colk[lane] = pik;
    __syncwarp();
    const double d = colk[k];
    if (!(d > 0.0)) not_pd = true;
    // one reciprocal per pivot, on the critical path of the whole knot: hardware seed (rel. error 2^-23) + two Newton
    // steps = four dependent multiply-adds, without the scaling / special-case code of the generic fp64 division
    const double rd = (fabs(d) > 2.2250738585072014e-308) ? rcp_newton(d) : 0.0;
    if (lane == 0) { s.D[k] = d; s.Dinv[k] = rd; }
    const double lik = pik * rd;
#pragma unroll
    for (int c = 1; c < n; ++c) {
      if (c <= k) continue;                                            // (constant trip counts: both loops unroll fully)
      p[c] -= lik * colk[c];                                           // P(i,c) -= l_ik P(c,k)
    }
    if (lane > k && lane < n) s.Ls[k * n + lane] = lik;
  }
  __syncwarp();
  return not_pd;
}

#ifdef RIC_PROF   // debug build only (tools/ric_prof.sh): per-phase cycles (work before / wait at each block barrier) of warp 0 and warp 7, summed over all blocks
__device__ unsigned long long ric_prof_sum[2][20];
#define RP_DECL long long rp_t = clock64(); long long rp_acc[20]; for (int q_ = 0; q_ < 20; ++q_) rp_acc[q_] = 0;
#define RP_MARK(p) { const long long now_ = clock64(); rp_acc[p] += now_ - rp_t; rp_t = now_; }
#define RP_SYNC(p) { RP_MARK(2 * (p)); __syncthreads(); { unsigned x_; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x_) : "r"((unsigned)__cvta_generic_to_shared(&s.perm[0])) : "memory"); rp_acc[19] += x_ & 0; } RP_MARK(2 * (p) + 1); }
#define RP_PRINT if (lane == 0 && (warp == 0 || warp == 7)) { for (int q_ = 0; q_ < 20; ++q_) atomicAdd(&ric_prof_sum[warp == 7][q_], (unsigned long long)rp_acc[q_]); }
#else
#define RP_DECL
#define RP_MARK(p)
#define RP_SYNC(p) __syncthreads();
#define RP_PRINT
#endif

template <bool STRUCT>
__global__ void __launch_bounds__(RIC_THREADS, 2)
k_backward(int N, const int* __restrict__ mask, const double* __restrict__ lambda, const double* __restrict__ A,
           const double* __restrict__ Bm, const double* __restrict__ lx, const double* __restrict__ lu,
           const double* __restrict__ lxx, const double* __restrict__ luu, double* __restrict__ K,
           double* __restrict__ kff, int* __restrict__ status, double h) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RiccatiSmem& s = *reinterpret_cast<RiccatiSmem*>(smem_raw);
  const int inst = blockIdx.x;
  if (mask && !mask[inst]) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const double lam = lambda[inst];
  const double* lxN = lx + ((size_t)inst * (N + 1) + N) * NX;
  const double* lxxN = lxx + ((size_t)inst * (N + 1) + N) * LXX_STRIDE;
  // [A_t | B_t] -> s.AB: two bulk copies issued by ONE thread (barrier s.mbar[1]); lx_t, lu_t, luu_t -> s.lq: 8-byte async copies
  // of the calling warp (their global strides are odd numbers of doubles), read in the epilogues of G2 / G3
  auto prefetch_ab = [&](int t) {
    if (lane == 0) {
      mbar_expect_tx(&s.mbar[1], (A_STRIDE + B_STRIDE) * 8);
      bulk_g2s(s.AB, A + ((size_t)inst * N + t) * A_STRIDE, A_STRIDE * 8, &s.mbar[1]);
      bulk_g2s(s.AB + A_STRIDE, Bm + ((size_t)inst * N + t) * B_STRIDE, B_STRIDE * 8, &s.mbar[1]);
    }
    const double* lxt = lx + ((size_t)inst * (N + 1) + t) * NX;
    const double* lut = lu + ((size_t)inst * N + t) * NU;
    const double* luut = luu + ((size_t)inst * N + t) * NU * NU;
    for (int i = lane; i < NX + NU + NU * NU; i += 32)
      cp_async8(&s.lq[i], i < NX ? lxt + i : (i < NX + NU ? lut + (i - NX) : luut + (i - NX - NU)));
  };
  // column c of [A|B] in s.AB (the pad double of A sits between the two blocks); row 51 of a column is the first entry of the next
  // one (or a pad): finite, and always multiplied by a zero of the other operand (V and W keep their zero pad row / column)
  auto abcol = [&](int c) -> const double* { return s.AB + c * NX + (c >= NX ? 1 : 0); };
  // zero pads (never written afterwards) and the terminal value function
  for (int i = tid; i < LDX * LDX; i += nt) s.V[i] = 0.0;
  for (int i = tid; i < LDX * NXU; i += nt) s.AB[i] = 0.0;
  if (tid == 0) { mbar_init(&s.mbar[0], 1); mbar_init(&s.mbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  for (int i = tid; i < LDU * 56; i += nt) s.Kt[i] = 0.0;
  for (int i = tid; i < LDU * LDU; i += nt) { s.Quu[i] = 0.0; s.Li[i] = 0.0; }
  fence_proxy_async();   // the zero fill above (generic proxy) is ordered before the copy engine's writes
  __syncthreads();
  if (warp == 7) prefetch_ab(N - 1);
  for (int i = tid; i < LDX; i += nt) s.Vx[i] = (i < NX) ? lxN[i] : 0.0;
  for (int i = tid; i < NX * NX; i += nt) {   // lxx holds its LOWER triangle (k_cost_quadratics): mirrored on load
    const int c = i / NX, r = i - c * NX;
    s.V[c * LDX + r] = lxxN[min(c, r) * NX + max(c, r)];
  }
  bool nonfinite = false;
  double* const G = s.W + RIC_G_OFF;
  // Operand fetches are branch-free: out-of-range rows / columns are redirected to zero pads (a divergent branch around a
  // fragment load doubled the time of every tile that straddles an edge). zc = the pad column of V, 52 zeros at all times.
  const double* const zc = &s.V[NX * LDX];
  double* const Lpre = s.W + RIC_LXX_OFF;
  RP_DECL
  for (int t = N - 1; t >= 0; --t) {
    const double* lxt = s.lq;
    const double* lut = s.lq + NX;
    const double* luut = s.lq + NX + NU;
    cp_async_commit_wait_all();                     // (warp 7: lx, lu, luu of this knot)
    mbar_wait(&s.mbar[1], (N - 1 - t) & 1);         // [A_t | B_t] have landed
    RP_SYNC(0)
    // ---- G1b: W(:, 48..71) = Vxx [A|B](:, 48..71) — the B block, which Quu needs first (warps 0..6: one 8-row
    //      strip each) ----
    auto w_strip = [&](auto nt_tag, int n0) {
      constexpr int NT = decltype(nt_tag)::value;
      if constexpr (STRUCT) {
        mma_strip_store<NT>(8, 8 * warp, n0,
                            [&](int r, int k) {   // (V E)(r, k): column 51 of V is the zero pad
                              const int rr = min(r, LDX - 1);
                              return fma(h, s.V[ric_row_q(k) * LDX + rr], s.V[ric_row_v(k) * LDX + rr]);
                            },
                            [&](int k, int c) { return (c < NXU ? abcol(c) : zc)[ric_row_v(k)]; },
                            [&](int r, int c, double v) {
                              if (r < LDX && c < NXU) s.W[c * LDX + r] = ric_unit_row(c) ? v + s.V[c * LDX + r] : v;   // + (V P)(r, c)
                            });
      } else {
        mma_strip_store<NT>(13, 8 * warp, n0,
                            [&](int r, int k) { return s.V[k * LDX + min(r, LDX - 1)]; },          // pad row 51 of V is zero
                            [&](int k, int c) { return (c < NXU ? abcol(c) : zc)[k]; },
                            [&](int r, int c, double v) { if (r < LDX && c < NXU) s.W[c * LDX + r] = v; });   // row 51 of W = 0
      }
    };
    if (warp < 7) w_strip(std::integral_constant<int, 3>(), 48);
    RP_SYNC(1)
    // ---- G3: [Quu | B'Vx] = B' [W_B | Vx] + luu + lam I, 9 tiles over the 8 warps; the spare column 19 of the last
    //      tile column carries Vx and yields Qu = lu + B'Vx ----
    {
      auto fa = [&](int r, int k) { return abcol(NX + min(r, NU - 1))[STRUCT ? ric_row_v(k) : k]; };
      auto fb = [&](int k, int c) {
        const double* col = c < NU ? s.W + (NX + c) * LDX : (c == NU ? s.Vx : zc);
        if constexpr (STRUCT) return fma(h, col[ric_row_q(k)], col[ric_row_v(k)]);   // (E'[W_B | Vx])(k, c); W row 51, Vx[51] are zero
        else return col[k];
      };
      // one tile per warp: the 6 tiles of the lower triangle (mirrored into the upper one — LLT / LDLT only read one
      // triangle, and B'VxxB is symmetric up to rounding) and the two remaining tiles of the column that carries Qu
      const int mi = warp < 6 ? (warp < 1 ? 0 : (warp < 3 ? 1 : 2)) : warp - 6;
      const int nj = warp < 6 ? (warp < 1 ? 0 : (warp < 3 ? warp - 1 : warp - 3)) : 2;
      double c0, c1;
      mma_tile_split<STRUCT ? 8 : 13, 4>(8 * mi, 8 * nj, fa, fb, c0, c1);
      const int r = 8 * mi + g;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = 8 * nj + 2 * t4 + q;
        const double v = q ? c1 : c0;
        if (r >= NU) continue;
        if (c < NU) {
          if (warp < 6) {
            if (c <= r) {
              const double e = v + luut[c * NU + r] + ((c == r) ? lam : 0.0);
              s.Quu[c * LDU + r] = e;
              s.Quu[r * LDU + c] = e;
            }
          }
        } else if (c == NU) s.Qu[r] = lut[r] + v;
      }
    }
    RP_SYNC(2)
    // ---- warps 0..6: G1a: W(:, 0..47) = Vxx A(:, 0..47), then G2: [Qxx | Qxu | A'Vx] = A' [W | Vx] (Qxx -> s.V,
    //      Qxu -> s.Qxu, the spare column 70 of the last tile carries Vx and yields Qx = lx + A'Vx)
    //      | warp 7: pivoted LDL^T of Quu, then N = L^-1: the sequential section of the knot, beside both contractions
    //      (one barrier for the whole phase) ----
    if (warp < 7) {
      w_strip(std::integral_constant<int, 6>(), 0);
      asm volatile("bar.sync 1, 224;" ::: "memory");   // W complete (the seven contraction warps only)
      // G2 tiles: only the lower triangle of Qxx is ever read again (G5 takes r >= c and mirrors), so of the 7 x 9 tiles
      // of A' [W | Vx] 48 remain: every warp keeps the column tiles 6..8 of its own row strip (Qxx columns 48..50, Qxu,
      // Qx) and takes 4 of the 27 other lower tiles — rows w and 6 - w hold 8 of them together, split between the two
      // warps (warp 6 has 3 and repeats one without storing it). 7 DMMA per k step instead of 9.
      {
        const int par = 6 - warp;
        int tc[7];
        bool own[7];
#pragma unroll
        for (int q = 0; q < 3; ++q) { tc[q] = 6 + q; own[q] = true; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (warp <= 3) { own[3 + q] = q <= warp; tc[3 + q] = own[3 + q] ? q : q - warp - 1; }
          else { own[3 + q] = true; tc[3 + q] = min(warp - 3 + q, 5); }
        }
        const double* bp[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int c = 8 * tc[q] + g;
          bp[q] = (c < NXU ? s.W + c * LDX : (c == NXU ? s.Vx : zc)) + (STRUCT ? 0 : t4);
        }
        const double* const ap1 = abcol(8 * warp + g) + (STRUCT ? 0 : t4);   // A'(r,k) = A(k,r); rows 51..55: discarded garbage
        const double* const ap2 = abcol(8 * par + g) + (STRUCT ? 0 : t4);
        double acc[7][2];
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[q][0] = acc[q][1] = 0.0;
        if constexpr (STRUCT) {
#pragma unroll 1
          for (int ks = 0; ks < 8; ++ks) {
            const int rv = ric_row_v(4 * ks + t4), rq = ric_row_q(4 * ks + t4);
            const double a1 = ap1[rv], a2 = ap2[rv];                                   // R_A'(r, k)
#pragma unroll
            for (int q = 0; q < 7; ++q)                                                // (E'[W | Vx])(k, c): W row 51, Vx[51], zc are zero
              dmma884(acc[q][0], acc[q][1], own[q] ? a1 : a2, fma(h, bp[q][rq], bp[q][rv]));
          }
        } else {
#pragma unroll 1
          for (int ks = 0; ks < 13; ++ks) {
            const double a1 = ap1[4 * ks], a2 = ap2[4 * ks];
#pragma unroll
            for (int q = 0; q < 7; ++q) dmma884(acc[q][0], acc[q][1], own[q] ? a1 : a2, bp[q][4 * ks]);
          }
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          if (q == 6 && warp == 6) continue;                   // the repeated tile
          const int r = 8 * (own[q] ? warp : par) + g;
          if (r >= NX) continue;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * tc[q] + 2 * t4 + e;
            double v = acc[q][e];
            if (STRUCT && ric_unit_row(r)) v += c < NXU ? s.W[c * LDX + r] : (c == NXU ? s.Vx[r] : 0.0);   // + (P'[W | Vx])(r, c)
            if (c < NX) s.V[c * LDX + r] = v;
            else if (c < NXU) s.Qxu[(c - NX) * LDX + r] = v;
            else if (c == NXU) s.Qx[r] = lxt[r] + v;
          }
        }
      }
    } else {
      if (quu_ldlt(s)) {          // Eigen::LLT failed: Quu += 1e-4 I once, no re-check (quirk Q9), refactor
        for (int i = lane; i < NU; i += 32) s.Quu[i * LDU + i] += 1e-4;
        __syncwarp();
        quu_ldlt(s);
      }
      // N = L^-1 (unit lower; lane j owns column j, 171 multiply-adds). With N explicit the 52 pairs of triangular solves
      // of the knot become two small tensor-core contractions instead of 36 dependent shuffle / multiply-add steps per
      // right-hand side. (Computing N row by row inside the factorisation loop was measured slower.)
      // Right-looking substitution: once N(m, j) is final every later row takes its update at once (independent
      // multiply-adds), so the dependent chain of the whole inverse is 19 operations instead of one dot product per row.
      if (lane < LDU) {
        double x[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < NU; ++m) {
          const double xm = x[m];
          s.Li[m * LDU + lane] = xm;
#pragma unroll
          for (int i = m + 1; i < NU; ++i) x[i] -= s.Ls[m * NU + i] * xm;
        }
      }
    }
    fence_proxy_async();   // this phase's generic accesses to s.AB / s.W are ordered before the copy engine overwrites them
    RP_SYNC(7)
    // ---- K = -Quu^-1 Qxu', k = -Quu^-1 Qu with P Quu P' = L D L':  X = P' N' D^-1 N (P R), R = [Qxu' | Qu] (19 x 52).
    //      Each of the warps 0..6 owns one 8-column tile of R: Y = N (P R), Z = D^-1 Y parked in its own columns of
    //      s.Kt, X = N' Z; no exchange between warps ----
    if (warp < 7) {
      const int n0 = 8 * warp;
      const int nrhs = n0 + g;                                 // column of R = [Qxu' | Qu] this lane feeds: row pk of it
      const double* const rsrc = nrhs < NX ? s.Qxu + nrhs : s.Qu;
      const int rstride = nrhs < NX ? LDX : 1;
      double acc[3][2];
#pragma unroll
      for (int mi = 0; mi < 3; ++mi) acc[mi][0] = acc[mi][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const int k = 4 * ks + t4;
        const int pk = s.perm[min(k, NU - 1)];
        const double bv = rsrc[pk * rstride];
        const double b = (k < NU && nrhs <= NX) ? bv : 0.0;
#pragma unroll
        for (int mi = 0; mi < 3; ++mi) {
          const int r = min(8 * mi + g, LDU - 1);             // pad row 19 of Li is zero
          dmma884(acc[mi][0], acc[mi][1], s.Li[r * LDU + k], b);
        }
      }
#pragma unroll
      for (int mi = 0; mi < 3; ++mi) {
        const int r = 8 * mi + g;
        if (r < LDU) {
          const double di = r < NU ? s.Dinv[r] : 0.0;     // (0 for a vanishing pivot, as Eigen::LDLT::solve does)
          s.Kt[(n0 + 2 * t4) * LDU + r] = di * acc[mi][0];
          s.Kt[(n0 + 2 * t4 + 1) * LDU + r] = di * acc[mi][1];
        }
      }
      __syncwarp();
#pragma unroll
      for (int mi = 0; mi < 3; ++mi) acc[mi][0] = acc[mi][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const int k = 4 * ks + t4;
        const double b = s.Kt[(n0 + g) * LDU + k];
#pragma unroll
        for (int mi = 0; mi < 3; ++mi) {
          const int r = min(8 * mi + g, LDU - 1);             // pad column 19 of Li is zero
          dmma884(acc[mi][0], acc[mi][1], s.Li[k * LDU + r], b);   // N'(r, k) = N(k, r)
        }
      }
      __syncwarp();   // every lane has read Z before its columns are overwritten with the gains
#pragma unroll
      for (int mi = 0; mi < 3; ++mi) {
        const int r = 8 * mi + g;
        if (r < NU) {
          const int pr = s.perm[r];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int n = n0 + 2 * t4 + q;
            if (n <= NX) {
              const double v = -acc[mi][q];
              if (!isfinite(v)) nonfinite = true;
              s.Kt[n * LDU + pr] = v;
            }
          }
        }
      }
      __syncwarp();
      // ---- G4: G = Quu K + 2 Qxu', the strip of the SAME 8 columns this warp has just solved for (no barrier between the
      //      solves and G4); column 51 of the last strip is kff, so its product is Quu k, which the Vx update needs ----
      auto fa = [&](int r, int k) { return s.Quu[k * LDU + min(r, LDU - 1)]; };   // pad row 19 of Quu is zero
      auto fb = [&](int k, int c) { return s.Kt[c * LDU + k]; };
#pragma unroll
      for (int mi = 0; mi < 3; ++mi)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int r = 8 * mi + g, c = n0 + 2 * t4 + q;
          const double qv = s.Qxu[min(r, NU - 1) * LDX + min(c, NX - 1)];
          acc[mi][q] = (r < NU && c < NX) ? 2.0 * qv : 0.0;
        }
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {       // the three row tiles are independent chains
        const int k = 4 * ks + t4;
        const double b = fb(k, n0 + g);
#pragma unroll
        for (int mi = 0; mi < 3; ++mi) dmma884(acc[mi][0], acc[mi][1], fa(8 * mi + g, k), b);
      }
#pragma unroll
      for (int mi = 0; mi < 3; ++mi) {
        const int r = 8 * mi + g;
        if (r < LDU) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = n0 + 2 * t4 + q;
            if (c < NX) G[c * LDU + r] = acc[mi][q];
            else if (c == NX && r < NU) s.tmp[r] = acc[mi][q];
          }
        }
      }
    } else {
      // warp 7 has nothing to solve: it starts the copies — s.AB, s.W and s.lq are free since the barrier above: this knot's lxx
      // (G5 adds it; dense, its lower triangle is what k_cost_quadratics wrote) and the next knot's [A|B], lx, lu, luu
      if (lane == 0) {
        mbar_expect_tx(&s.mbar[0], LXX_STRIDE * 8);
        bulk_g2s(Lpre, lxx + ((size_t)inst * (N + 1) + t) * LXX_STRIDE, LXX_STRIDE * 8, &s.mbar[0]);
      }
      if (t > 0) prefetch_ab(t - 1);
    }
    RP_SYNC(4)
    {   // gains to global
      double* Kt_g = K + ((size_t)inst * N + t) * NU * NX;
      double* kf_g = kff + ((size_t)inst * N + t) * NU;
      for (int i = tid; i < NU * NX; i += nt) { const int c = i / NU, r = i - c * NU; Kt_g[i] = s.Kt[c * LDU + r]; }
      for (int i = tid; i < NU; i += nt) kf_g[i] = s.Kt[NX * LDU + i];
    }
    // ---- G5: Vxx = lxx + Qxx + K' G, lower triangle only, mirrored in place into s.V (warps 0..6: 4 of the 28 lower
    //      tiles each, split like the G2 tiles) | warp 7: Vx = Qx + K'(Quu k) + K'Qu + Qxu k.
    //      K'G = K'QuuK + 2 K'Qxu' is symmetric up to rounding (K = -S Qxu' with S = P'N'D^-1 N P symmetric by
    //      construction), so the reference's 0.5 (M + M') differs from the mirrored lower triangle by rounding only ----
    if (warp < 7) {
      mbar_wait(&s.mbar[0], (N - 1 - t) & 1);       // lxx_t has landed
      const int par = 6 - warp;
      int tc[4];
      bool own[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (warp <= 3) { own[q] = q <= warp; tc[q] = own[q] ? q : q - warp - 1; }
        else { own[q] = true; tc[q] = warp - 3 + q; }
      }
      double acc[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = 8 * (own[q] ? warp : par) + g, c = 8 * tc[q] + 2 * t4 + e;
          const int rc = min(r, NX - 1), cc = min(c, NX - 1);
          const double qv = s.V[cc * LDX + rc] + Lpre[cc * NX + rc];
          acc[q][e] = (r < NX && c <= r) ? qv : 0.0;
        }
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const int k = 4 * ks + t4;
        const double a1 = s.Kt[(8 * warp + g) * LDU + k];     // K'(r,k) = K(k,r); rows 51..55: kff / zeros, discarded
        const double a2 = s.Kt[(8 * par + g) * LDU + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = 8 * tc[q] + g;
          const double gv = G[c * LDU + k];                   // c = 51..55: reads the lxx prefetch behind G, discarded
          dmma884(acc[q][0], acc[q][1], own[q] ? a1 : a2, c < NX ? gv : 0.0);
        }
      }
      __syncwarp();   // the (discarded) clamped / upper-triangle reads of the accumulator init touch elements other lanes of this warp write below
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = 8 * (own[q] ? warp : par) + g, c = 8 * tc[q] + 2 * t4 + e;
          if (r < NX && c <= r) { s.V[c * LDX + r] = acc[q][e]; s.V[r * LDX + c] = acc[q][e]; }
        }
    } else {
      // both rows of a lane (i and i + 32) and the even / odd halves of each sum run as 12 independent chains
      const int i0 = lane, i1 = min(lane + 32, NX - 1);
      double a[2][3][2];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int m = 0; m < 3; ++m) a[h2][m][0] = a[h2][m][1] = 0.0;
#pragma unroll
      for (int l = 0; l < NU; ++l) {
        const double tl = s.tmp[l], ql = s.Qu[l], kl = s.Kt[NX * LDU + l];
        const double k0 = s.Kt[i0 * LDU + l], k1 = s.Kt[i1 * LDU + l];
        a[0][0][l & 1] += k0 * tl; a[0][1][l & 1] += k0 * ql; a[0][2][l & 1] += s.Qxu[l * LDX + i0] * kl;
        a[1][0][l & 1] += k1 * tl; a[1][1][l & 1] += k1 * ql; a[1][2][l & 1] += s.Qxu[l * LDX + i1] * kl;
      }
      s.Vx[i0] = s.Qx[i0] + (a[0][0][0] + a[0][0][1]) + (a[0][1][0] + a[0][1][1]) + (a[0][2][0] + a[0][2][1]);
      if (lane + 32 < NX) s.Vx[i1] = s.Qx[i1] + (a[1][0][0] + a[1][0][1]) + (a[1][1][0] + a[1][1][1]) + (a[1][2][0] + a[1][2][1]);
    }
    // (the __syncthreads at the top of the next iteration orders these writes before G1)
  }
  RP_PRINT
  if (nonfinite && status) status[inst] = 1;
}

}  // namespace h1
