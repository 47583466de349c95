// Batched analytic linearization in two steps (iLQR::computeLinearization, /root/reference/src/ilqr/ilqr.cpp:126-131;
// replaces RobotUtils::linearizeDynamicsFD, /root/reference/src/common/robot_utils.cpp:120-160, by exact tangents):
//
//   1. k_linearize_tangents<CLS>: the inverse-dynamics tangents t_c = -dg/d(direction c) of the 49 state directions
//      that are not trivial (h1_lin_dirs.cuh), one THREAD per (knot, direction), direction-uniform warps over 32
//      knots per CTA. The 25 entries of t_c are parked in the first 25 rows of column c of A_k (the three base
//      rotation tangents in columns 3..5); no triangular solve, no integrator tangent in this kernel.
//   2. k_linearize_finish: one WARP per knot turns the parked tangents into [A_k | B_k]:
//        N = L^-1 (unit lower, tree sparse) from the factor Mhat = L' D L kept by the rollout,
//        Mhat^-1 = N D^-1 N'                      (fp64 tensor cores, only the non-zero k blocks of N),
//        Adot = Mhat^-1 T   (25 x 25 x 51)        (fp64 tensor cores; the control columns are columns of Mhat^-1),
//        [A | B] = integrator tangent of Adot     (semi-implicit Euler + quaternion exponential Jacobians),
//      written as contiguous columns. All 70 triangular solves of a knot become one dense contraction.
#pragma once
#include "h1_lin_dirs.cuh"

namespace h1 {

#if defined(__CUDACC__)

constexpr int LINT_WARPS = 8, LINT_KNOTS = 32, LINT_THREADS = LINT_WARPS * 32;
__host__ __device__ constexpr int lint_nitems(int cls) { return cls == 0 ? NB - 1 : cls == 1 ? 3 + NB - 1 : 8; }
constexpr int LINT_SC = 2 * (NB - 1) + 1;   // sin / cos of the 19 hinges per knot (odd stride)
constexpr size_t LINT_SMEM_DOUBLES = (size_t)LINT_KNOTS * (NV + NX + LINT_SC) + (size_t)LINT_WARPS * LINT_KNOTS * NV;

H1_DEV void quat_rot_map_col(const double* __restrict__ qraw, int i, double* g3) {
  Dual q[4], qn[4], Rd[9];
  for (int j = 0; j < 4; ++j) q[j] = Dual(qraw[j], j == i ? 1.0 : 0.0);
  quat_normalize(q, qn);
  quat_to_mat(qn, Rd);
  auto w = [&](int r, int c) { return Rd[3 * r].d * Rd[3 * c].v + Rd[3 * r + 1].d * Rd[3 * c + 1].v + Rd[3 * r + 2].d * Rd[3 * c + 2].v; };
  g3[0] = w(2, 1); g3[1] = w(0, 2); g3[2] = w(1, 0);
}

// slot -> global knot id (instance * N + t) through the optional compact active list; -1: nothing to do
__device__ __forceinline__ long lin_knot_id(long slot, long nslots, int N, const int* __restrict__ active,
                                            const int* __restrict__ list) {
  if (slot >= nslots) return -1;
  if (list) return (long)list[slot / N] * N + slot % N;
  return (!active || active[slot / N]) ? slot : -1;
}

// CLS 0: hinge angles (19 items, subtree walks on dual kinematics)
//     1: base angular velocity (3, every body) + hinge rates (19, subtree walks), plain kinematics
//     2: rigid directions: base rotations (3), z, base linear velocity (3); plus one item per CTA that computes the
//        per-knot Jacobians of the quaternion update and the quaternion -> rotation-vector map for step 2 (parked in
//        the unused rows 25.. of columns 0 and 1 of A_k)
//     3: ALL of the above in one launch, the 49 items of a knot group in the order DynModel::tan_order (class by class, costliest
//        first: model_tables.cpp has the measurements). One launch per class left the eight warps of a CTA idle at the end of every
//        group, and staged the group's states / accelerations / sines three times.
#ifdef LINT_PROF   // debug build only (tools/lint_prof.py): cycles per item code, per-warp busy time and CTA makespan, summed over all CTAs
__device__ unsigned long long lint_prof_item[128], lint_prof_busy, lint_prof_span, lint_prof_stage;
#endif
template <int CLS>
__global__ void __launch_bounds__(LINT_THREADS)
k_linearize_tangents(const DynModel* gmd, long nknots, int N, const int* __restrict__ active,
                     const int* __restrict__ list, const int* __restrict__ list_count,
                     const double* __restrict__ xbar, const PrimalFactor* __restrict__ pf_g, double* __restrict__ A) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ long kid[LINT_KNOTS];
  __shared__ int next_item;
  const long knot0 = (long)blockIdx.x * LINT_KNOTS;
  if (list) nknots = min(nknots, (long)(*list_count) * N);
  if (knot0 >= nknots) return;
  const int nk = (int)min((long)LINT_KNOTS, nknots - knot0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < LINT_KNOTS) kid[tid] = lin_knot_id(knot0 + tid, nknots, N, active, list);
  if (tid == 0) next_item = 0;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);    // (has a __syncthreads)
  bool any = false;
  for (int k = 0; k < nk; ++k) any |= kid[k] >= 0;
  if (!any) return;
  double* as = reinterpret_cast<double*>(p);          // [LINT_KNOTS][NV] primal accelerations
  double* xs = as + LINT_KNOTS * NV;                  // [LINT_KNOTS][NX]
  double* scs = xs + LINT_KNOTS * NX;                // [LINT_KNOTS][LINT_SC] sin / cos of the hinge angles, shared by all directions
  double* tile = scs + LINT_KNOTS * LINT_SC + warp * LINT_KNOTS * NV;   // this warp's [LINT_KNOTS][NV]
  for (int i = tid; i < nk * NV; i += LINT_THREADS) {
    const int k = i / NV, j = i - k * NV;
    const long id = kid[k];
    if (id >= 0) as[i] = pf_g[id].a[j];
  }
  for (int i = tid; i < nk * NX; i += LINT_THREADS) {
    const int k = i / NX, j = i - k * NX;
    const long id = kid[k];
    if (id < 0) continue;
    const long inst = id / N;
    xs[i] = xbar[((size_t)inst * (N + 1) + (id - inst * N)) * NX + j];
  }
  __syncthreads();
  for (int i = tid; i < nk * (NB - 1); i += LINT_THREADS) {
    const int k = i / (NB - 1), b = i - k * (NB - 1);
    if (kid[k] >= 0) sincos_t(xs[k * NX + 7 + b], &scs[k * LINT_SC + 2 * b], &scs[k * LINT_SC + 2 * b + 1]);
  }
  __syncthreads();
#ifdef LINT_PROF
  const long long lp_t0 = clock64();
  long long lp_busy = 0;
#endif
  const bool ok = lane < nk && kid[lane] >= 0;
  const double* x = xs + lane * NX;
  const double* a = as + lane * NV;
  const double* sc = scs + lane * LINT_SC;
  double* tl = tile + lane * NV;
  while (true) {
    int it = 0;
    if (lane == 0) it = atomicAdd(&next_item, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
#ifdef LINT_PROF
    const long long lp_i0 = clock64();
    const int lp_it = it;
#endif
    int cls = CLS;
    if (CLS == 3) {
      if (it >= md->n_tan_items) break;
      const int code = md->tan_order[it];
      cls = code >> 5; it = code & 31;
    } else if (it >= lint_nitems(CLS)) break;
    int col;                                          // column of A_k the tangent is parked in
    if (cls == 2 && it == 3) {                        // quaternion Jacobians J (28) and rotation map G (12) of each knot
      if (ok) {
        double* park = A + (size_t)kid[lane] * A_STRIDE + NV;       // rows 25..50 of column 0, then rows 25.. of column 1
        for (int d = 0; d < QJ_DIRS; ++d) {
          double J4[4];
          quat_step_jac_dir(*md, x, a, d, J4);
          for (int r = 0; r < 4; ++r) { const int i = 4 * d + r; park[i < 26 ? i : NX + i - 26] = J4[r]; }
        }
        for (int i = 0; i < 4; ++i) {
          double g3[3];
          quat_rot_map_col(x + 3, i, g3);
          for (int r = 0; r < 3; ++r) park[NX + 2 + 3 * i + r] = g3[r];
        }
      }
      continue;
    }
    if (ok) {
      double* const tv = tl;   // straight into this lane's row of the warp's tile (shared memory): a local array here is 200 B more
                               // of per-thread stack, whose write-back traffic is what this kernel's DRAM writes mostly are
      if (cls == 0) {
        col = 6 + md->dir_order[it];                  // hinges by decreasing subtree size
        id_tangent_sub<Dual, Dual>(*md, x, a, col, col - 6, tv, sc);
      } else if (cls == 1) {
        if (it < 3) { col = NQ + 3 + it; id_tangent_seq<double, Dual>(*md, x, a, col, tv, sc); }
        else { col = NQ + 5 + md->dir_order[it - 3]; id_tangent_sub<double, Dual>(*md, x, a, col, col - NQ - 5, tv, sc); }
      } else {
        if (it < 3) { col = 3 + it; id_tangent_rot(*md, x, a, it, tv, sc); }
        else { col = it == 4 ? 2 : NQ + it - 5; id_tangent_rigid(*md, x, a, col, tv, sc); }
      }
    }
    if (cls == 0) col = 6 + md->dir_order[it];
    else if (cls == 1) col = it < 3 ? NQ + 3 + it : NQ + 5 + md->dir_order[it - 3];
    else col = it < 3 ? 3 + it : (it == 4 ? 2 : NQ + it - 5);
    __syncwarp();
    {
      double* dst0 = A + (size_t)col * NX;
      int k = 0, j = lane;
      while (j >= NV) { j -= NV; ++k; }
#pragma unroll 1
      for (int e = lane; e < nk * NV; e += 32) {
        const long id = kid[k];
        if (id >= 0) dst0[(size_t)id * A_STRIDE + j] = tile[e];
        j += 32;
        while (j >= NV) { j -= NV; ++k; }
      }
    }
    __syncwarp();
#ifdef LINT_PROF
    { const long long d = clock64() - lp_i0; lp_busy += d; if (lane == 0) atomicAdd(&lint_prof_item[CLS == 3 ? md->tan_order[lp_it] : (CLS << 5 | lp_it)], (unsigned long long)d); }
#endif
  }
#ifdef LINT_PROF
  if (lane == 0) atomicAdd(&lint_prof_busy, (unsigned long long)lp_busy);
  __syncthreads();
  if (tid == 0) { atomicAdd(&lint_prof_span, (unsigned long long)(clock64() - lp_t0)); }
#endif
}

// ---- step 2 ----
#ifdef LINF_PROF   // debug build only (tools/linf_prof.py): cycles of a warp between the numbered steps of k_linearize_finish, summed over all knots
__device__ unsigned long long linf_prof_sum[8];
#define LF_MARK(p) { const long long n_ = clock64(); if (lane == 0) atomicAdd(&linf_prof_sum[p], (unsigned long long)(n_ - lf_t)); lf_t = n_; }
#else
#define LF_MARK(p)
#endif
constexpr int LINF_WARPS = 5, LINF_THREADS = LINF_WARPS * 32;   // 22.5 KB of shared memory per warp: two CTAs = 10 warps per SM
constexpr int LDF = 28;                 // leading dimension of the 25-row operands: = 4 (mod 8) doubles, k padded to 28
struct LinFinishWarp {
  double Nm[32 * LDF];                  // L^-1, row-major (rows / columns >= 25 are zero); then Mhat^-1 in place
  double Fs[sizeof(PrimalFactor) / sizeof(double) + 3];   // factor staging (PrimalFactor)
  double T[NX * LDF];                   // tangents, column c at T[c * LDF]; then Adot in place
  double Dinv[LDF];
  double x[NX + 1], a[NV + 1], J[QJ_DIRS * 4], G[12], umask[NU + 1];
};

__device__ __forceinline__ void lf_cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}

template <bool H1TREE>
__global__ void __launch_bounds__(LINF_THREADS, 2)
k_linearize_finish(const DynModel* __restrict__ md, long nknots, int N, const int* __restrict__ active,
                   const int* __restrict__ list, const int* __restrict__ list_count,
                   const double* __restrict__ xbar, const double* __restrict__ ubar,
                   const PrimalFactor* __restrict__ pf_g, double* __restrict__ A, double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  LinFinishWarp& W = reinterpret_cast<LinFinishWarp*>(smem)[warp];
  if (list) nknots = min(nknots, (long)(*list_count) * N);
  const long id = lin_knot_id((long)blockIdx.x * LINF_WARPS + warp, nknots, N, active, list);
  if (id < 0) return;                                  // (no block-level barrier below)
  const double h = md->h;
  double* Ak = A + (size_t)id * A_STRIDE;
  double* Bk = Bm + (size_t)id * B_STRIDE;
#ifdef LINF_PROF
  long long lf_t = clock64();
#endif
  // ---- 1. stage with asynchronous copies in two groups: (A) factor + state, needed at once; (B) the parked tangents,
  //      needed only by the contraction of step 4 — their memory latency hides behind N = L^-1 and Mhat^-1.
  //      Columns 0, 1, 6 and the pad rows of T are zero ----
  {
    const double* src = reinterpret_cast<const double*>(pf_g + id);
    for (int i = lane; i < (int)(sizeof(PrimalFactor) / sizeof(double)); i += 32) lf_cp_async8(&W.Fs[i], src + i);
    const long inst = id / N;
    const double* xg = xbar + ((size_t)inst * (N + 1) + (id - inst * N)) * NX;
    for (int i = lane; i < NX; i += 32) lf_cp_async8(&W.x[i], xg + i);
    if (lane < 26) lf_cp_async8(&W.J[lane], Ak + NV + lane);                      // parked by k_linearize_tangents<2>
    if (lane < 14) lf_cp_async8(lane < 2 ? &W.J[26 + lane] : &W.G[lane - 2], Ak + NX + NV + lane);
    if (lane < NU) lf_cp_async8(&W.umask[lane], ubar + (size_t)id * NU + lane);   // (turned into the clamp mask below)
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (lane < NV) {
#pragma unroll 4
      for (int c = 2; c < NX; ++c)
        if (c != 6) lf_cp_async8(&W.T[c * LDF + lane], Ak + c * NX + lane);
      W.T[lane] = 0.0; W.T[LDF + lane] = 0.0; W.T[6 * LDF + lane] = 0.0;
    } else if (lane < LDF) {
      for (int c = 0; c < NX; ++c) W.T[c * LDF + lane] = 0.0;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int i = lane; i < 32 * LDF; i += 32) W.Nm[i] = 0.0;
    LF_MARK(0)   // issue of the copies + zero fill
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");            // group A has landed
    LF_MARK(1)   // wait for the factor
    if (lane < NU) {
      const double uj = W.umask[lane];
      W.umask[lane] = (uj < md->ctrl_lo[lane] || uj > md->ctrl_hi[lane]) ? 0.0 : 1.0;   // clamped torque: no sensitivity
    }
  }
  __syncwarp();
  // ---- 2. per-knot small quantities and N = L^-1: lane c owns column c ----
  {
    const PrimalFactor* pf = reinterpret_cast<const PrimalFactor*>(W.Fs);
    if (lane < LDF) W.Dinv[lane] = lane < NV ? 1.0 / pf->D[lane] : 0.0;
    if (lane < NV) W.a[lane] = pf->a[lane];
    __syncwarp();
    if (lane < LDF) {
      const int c = lane;
      if (H1TREE) {
        double col[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          double v = (c == k) ? 1.0 : 0.0;
#pragma unroll
          for (int s = 0; s < MAXSLOT - 1; ++s)
            if (s < h1_nlist(k) - 1) v -= pf->Lm[k][s] * col[h1_anc(k, s)];
          col[k] = v;
          W.Nm[k * LDF + c] = v;
        }
      } else {
        for (int k = 0; k < NV; ++k) {
          double v = (c == k) ? 1.0 : 0.0;
          const int n = md->nlist[k];
          for (int s = 0; s < n - 1; ++s) v -= pf->Lm[k][s] * W.Nm[md->alist[k][s] * LDF + c];
          W.Nm[k * LDF + c] = v;
        }
      }
    }
  }
  __syncwarp();
  LF_MARK(2)   // N = L^-1
  // ---- 3. Mhat^-1 = (N D^-1) N' : tile (mi, nj) only needs k < 8 (min(mi, nj) + 1) (N is lower triangular). All 16
  //      tiles are accumulated in registers and then written over N ----
  {
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const int k = 4 * ks + t4;
      const double dk = W.Dinv[k];
      double nop[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) nop[q] = (ks < 2 * (q + 1)) ? W.Nm[(8 * q + g) * LDF + k] : 0.0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        if (ks >= 2 * (mi + 1)) continue;
        const double aop = nop[mi] * dk;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          if (ks >= 2 * (nj + 1)) continue;
          dmma884(acc[mi][nj][0], acc[mi][nj][1], aop, nop[nj]);
        }
      }
    }
    __syncwarp();   // every lane has read N
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj) {
        const int r = 8 * mi + g, c = 8 * nj + 2 * t4;
        if (c < LDF) { W.Nm[r * LDF + c] = acc[mi][nj][0]; W.Nm[r * LDF + c + 1] = acc[mi][nj][1]; }
      }
  }
  __syncwarp();
  LF_MARK(3)   // Mhat^-1
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");              // group B (the tangents) has landed
  LF_MARK(4)   // wait for the tangents
  __syncwarp();
  // the four raw-quaternion tangents = combinations of the three rotation tangents (parked in columns 3..5)
  if (lane < NV) {
    const double r0 = W.T[3 * LDF + lane], r1 = W.T[4 * LDF + lane], r2 = W.T[5 * LDF + lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) W.T[(3 + i) * LDF + lane] = W.G[3 * i] * r0 + W.G[3 * i + 1] * r1 + W.G[3 * i + 2] * r2;
  }
  __syncwarp();
  // ---- 4. Adot = Mhat^-1 T (25 x 28 x 51 in 4 x 7 tiles), written back over T ----
  {
    double acc[4][7][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 7; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
#pragma unroll 1
    for (int ks = 0; ks < 7; ++ks) {
      const int k = 4 * ks + t4;
      double aop[4], bop[7];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) aop[mi] = W.Nm[(8 * mi + g) * LDF + k];
#pragma unroll
      for (int nj = 0; nj < 7; ++nj) { const int n = 8 * nj + g; bop[nj] = n < NX ? W.T[n * LDF + k] : 0.0; }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nj = 0; nj < 7; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], aop[mi], bop[nj]);
    }
    __syncwarp();
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 7; ++nj) {
        const int r = 8 * mi + g;
        if (r >= LDF) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) { const int n = 8 * nj + 2 * t4 + q; if (n < NX) W.T[n * LDF + r] = acc[mi][nj][q]; }
      }
  }
  __syncwarp();
  LF_MARK(5)   // Adot = Mhat^-1 T
  // ---- 5. integrator tangent (state order: q(26) then v(25)). A lane owns output rows `lane` and `lane + 32` of all
  //      70 columns: row r reads Adot row j(r); position rows get the extra factor h and the unit entries of
  //      d q_next / d q and h d q_next / d v_next. The four quaternion rows follow from the staged Jacobians ----
  {
    const int r0 = lane, r1 = lane + 32;                        // r1 < NX  <=>  lane < 19 (all velocity rows)
    const bool quat0 = r0 >= 3 && r0 < 7, pos0 = r0 < NQ;
    const int j0 = r0 >= NQ ? r0 - NQ : (r0 < 3 ? r0 : (r0 >= 7 ? r0 - 1 : 3));
    const int j1 = r1 - NQ;
    const bool has1 = r1 < NX;
    const int cv0 = NQ + j0, cv1 = NQ + j1;                     // columns with a unit velocity entry in these rows
    {
      const double* ad = W.T;
      double* dst = Ak;
#pragma unroll 6
      for (int c = 0; c < NX; ++c, ad += LDF, dst += NX) {
        const double b = ((c == cv0) ? 1.0 : 0.0) + h * ad[j0];
        const double v0 = pos0 ? ((c == r0) ? 1.0 : 0.0) + h * b : b;
        if (!quat0) dst[r0] = v0;
        if (has1) dst[r1] = ((c == cv1) ? 1.0 : 0.0) + h * ad[j1];
      }
    }
    {
      double* dst = Bk;
#pragma unroll 4
      for (int j = 0; j < NU; ++j, dst += NX) {
        const double* ad = W.Nm + (6 + j) * LDF;                // (Mhat^-1 is symmetric: row = column)
        const double sc = h * W.umask[j];
        const double b = sc * ad[j0];
        if (!quat0) dst[r0] = pos0 ? h * b : b;
        if (has1) dst[r1] = sc * ad[j1];
      }
    }
#pragma unroll 1
    for (int e = lane; e < (NX + NU) * 4; e += 32) {            // quaternion rows: one (column, row) pair per lane
      const int c = e >> 2, q = e & 3;
      const bool isu = c >= NX;
      const double* ad = isu ? W.Nm + (6 + c - NX) * LDF : W.T + c * LDF;
      const double sc = isu ? h * W.umask[c - NX] : h;
      const double w0 = ((c == NQ + 3) ? 1.0 : 0.0) + sc * ad[3], w1 = ((c == NQ + 4) ? 1.0 : 0.0) + sc * ad[4],
                   w2 = ((c == NQ + 5) ? 1.0 : 0.0) + sc * ad[5];
      double v = W.J[16 + q] * w0 + W.J[20 + q] * w1 + W.J[24 + q] * w2;
      if (c >= 3 && c < 7) v += W.J[4 * (c - 3) + q];
      (isu ? Bk + (size_t)(c - NX) * NX : Ak + (size_t)c * NX)[3 + q] = v;
    }
  }
  LF_MARK(6)   // integrator tangent + stores
}

#endif  // __CUDACC__

}  // namespace h1
