// Kernels built on the warp-cooperative f_D: batched one-step dynamics, nominal rollout with on-device cost,
// forward-difference linearization, and kinematic queries.
#pragma once
#include "h1_cost_eval.cuh"
#include "h1_lin_dirs.cuh"

namespace h1 {

// Stage the (read-only) model into shared memory; returns the aligned start of the per-warp scratch area.
template <class Model>
__device__ __forceinline__ unsigned char* stage_model(unsigned char* smem, const Model* g, const Model** out) {
  Model* s = reinterpret_cast<Model*>(smem);
  const int nwords = sizeof(Model) / 4;
  const unsigned* src = reinterpret_cast<const unsigned*>(g);
  unsigned* dst = reinterpret_cast<unsigned*>(s);
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  *out = s;
  return smem + ((sizeof(Model) + 15) / 16) * 16;
}

// ---- x_next[i] = f_D(x[i], u[i]) : one warp per state (RobotUtils::rolloutOneStep / step) ----
__global__ void k_dyn_step(const DynModel* gmd, int n, const double* __restrict__ x, const double* __restrict__ u,
                           double* __restrict__ xn) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  dyn_step_warp(*md, w, x + (size_t)i * NX, u + (size_t)i * NU, xn + (size_t)i * NX);
}

// ---- bias forces, dynamics-model CoM, ankle positions and sole contact points of arbitrary states
//      (computeGravComp, loadReferences' per-row FK, contact-schedule generation) ----
__global__ void k_dyn_query(const DynModel* gmd, int n, const double* __restrict__ x, double* __restrict__ bias,
                            double* __restrict__ com, double* __restrict__ ee, double* __restrict__ sole) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  dyn_assemble_warp(*md, w, x + (size_t)i * NX, nullptr);
  if (bias && lane < NV) bias[(size_t)i * NV + lane] = w.biasv[lane];
  if (com && lane < 3) com[(size_t)i * 3 + lane] = w.com[lane];
  if (ee && lane < 6) ee[(size_t)i * 6 + lane] = w.q[lane % 3] + w.footr[lane / 3][lane % 3];
  if (sole && lane < NCPT)
    for (int c = 0; c < 3; ++c) sole[((size_t)i * NCPT + lane) * 3 + c] = w.q[c] + w.cp[lane][c];
}

// ---- nominal rollout xbar[t+1] = f_D(xbar[t], ubar[t]) with the trajectory cost as a by-product
//      (iLQR::forwardRolloutNominal + the baseline computeTotalCost of the following line search).
//      One warp per instance. t_begin > 0 rolls out only the tail (warm start: last knot). ----
__global__ void k_rollout(const DynModel* gmd, const H1Weights* gw, RefTable refs, int B, int N, int t_begin,
                          const int* __restrict__ active, const double* __restrict__ x0, double* __restrict__ xbar,
                          const double* __restrict__ ubar, double* __restrict__ cost_out,
                          PrimalFactor* __restrict__ pf_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (inst >= B) return;
  if (active && !active[inst]) return;
  const int lane = threadIdx.x & 31;
  double* xb = xbar + (size_t)inst * (N + 1) * NX;
  const double* ub = ubar + (size_t)inst * N * NU;
  if (x0) {
    for (int i = lane; i < NX; i += 32) xb[i] = x0[(size_t)inst * NX + i];
    __syncwarp();
  }
  const RefView r = refs.view(inst);
  double total = 0.0;
  for (int t = 0; t < N; ++t) {
    if (t >= t_begin) {
      if (pf_out) dyn_primal_factor_warp(*md, w, xb + t * NX, ub + t * NU, xb + (t + 1) * NX, pf_out[(size_t)inst * N + t]);
      else dyn_step_warp(*md, w, xb + t * NX, ub + t * NU, xb + (t + 1) * NX);
    } else if (cost_out) {
      dyn_assemble_warp(*md, w, xb + t * NX, ub + t * NU);
    }
    if (cost_out) total += knot_cost_warp(*md, w, *gw, r, t, ub + t * NU, false);
    __syncwarp();
  }
  if (cost_out) {
    dyn_assemble_warp(*md, w, xb + N * NX, nullptr);
    total += knot_cost_warp(*md, w, *gw, r, N, nullptr, true);
    if (lane == 0) cost_out[inst] = total;
  }
}

// ---- forward-difference linearization (iLQR::computeLinearization -> RobotUtils::linearizeDynamicsFD):
//      A[:,i] = (f(x + eps e_i, u) - f(x,u)) / eps, B[:,j] likewise; 1 + 51 + 19 evaluations per knot.
//      One CTA per (instance, knot); its warps share the 71 evaluations through a shared-memory table. ----
constexpr int LIN_WARPS = 6;
constexpr int LIN_EVALS = 1 + NX + NU;
__global__ void __launch_bounds__(LIN_WARPS * 32)
k_linearize_fd(const DynModel* gmd, int N, double eps, const int* __restrict__ active,
               const double* __restrict__ xbar, const double* __restrict__ ubar, double* __restrict__ A,
               double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int inst = blockIdx.x / N, t = blockIdx.x % N;
  if (active && !active[inst]) return;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp* ws = reinterpret_cast<DynWarp*>(p);
  double* fx = reinterpret_cast<double*>(p + sizeof(DynWarp) * LIN_WARPS);  // [LIN_EVALS][NX]
  double* xs = fx + LIN_EVALS * NX;                                         // [NX + NU] staged inputs
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const double* u = ubar + ((size_t)inst * N + t) * NU;
  for (int i = threadIdx.x; i < NX + NU; i += blockDim.x) xs[i] = (i < NX) ? x[i] : u[i - NX];
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  for (int e = warp; e < LIN_EVALS; e += LIN_WARPS)
    dyn_step_warp(*md, ws[warp], xs, xs + NX, fx + e * NX, e - 1, eps);
  __syncthreads();
  double* Ak = A + ((size_t)inst * N + t) * NX * NX;
  double* Bk = Bm + ((size_t)inst * N + t) * NX * NU;
  const double inv = 1.0 / eps;
  (void)inv;
  for (int i = threadIdx.x; i < NX * (NX + NU); i += blockDim.x) {
    const int col = i / NX, row = i - col * NX;
    const double v = (fx[(1 + col) * NX + row] - fx[row]) / eps;
    if (col < NX) Ak[i] = v; else Bk[i - NX * NX] = v;
  }
}

// ---- factorisation of Mhat at every knot of the current trajectory (granular API path; inside a solve the
//      nominal rollout produces the same factors as a by-product) ----
__global__ void k_primal_factor(const DynModel* gmd, int B, int N, const int* __restrict__ active,
                                const double* __restrict__ xbar, const double* __restrict__ ubar,
                                PrimalFactor* __restrict__ pf_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  DynWarp& w = reinterpret_cast<DynWarp*>(p)[threadIdx.x >> 5];
  const long k = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= (long)B * N) return;
  const int inst = (int)(k / N), t = (int)(k % N);
  if (active && !active[inst]) return;
  dyn_primal_factor_warp(*md, w, xbar + ((size_t)inst * (N + 1) + t) * NX, ubar + ((size_t)inst * N + t) * NU, nullptr,
                         pf_out[k]);
}

// ---- analytic linearization: A = d f_D/dx, B = d f_D/du exactly. Per knot, ONE factorisation of Mhat (taken
//      from the nominal rollout) serves all 70 tangent directions; each direction is a dual-number
//      inverse-dynamics pass + two sparse triangular solves (h1_dyn.cuh). One CTA per (instance, knot),
//      directions round-robin over its warps, columns written straight to A_k / B_k. ----
constexpr int LINA_WARPS = 4;
__global__ void __launch_bounds__(LINA_WARPS * 32, 3)
k_linearize_analytic(const DynModel* gmd, int N, const int* __restrict__ active, const double* __restrict__ xbar,
                     const double* __restrict__ ubar, const PrimalFactor* __restrict__ pf_g, double* __restrict__ A,
                     double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int inst = blockIdx.x / N, t = blockIdx.x % N;
  if (active && !active[inst]) return;
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);
  TanWarpT<Dual>* ws = reinterpret_cast<TanWarpT<Dual>*>(p);
  PrimalFactor* pf = reinterpret_cast<PrimalFactor*>(p + sizeof(TanWarpT<Dual>) * LINA_WARPS);
  double* xs = reinterpret_cast<double*>(pf + 1);
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const double* u = ubar + ((size_t)inst * N + t) * NU;
  for (int i = threadIdx.x; i < NX + NU; i += blockDim.x) xs[i] = (i < NX) ? x[i] : u[i - NX];
  {
    const double* src = reinterpret_cast<const double*>(pf_g + ((size_t)inst * N + t));
    double* dst = reinterpret_cast<double*>(pf);
    for (int i = threadIdx.x; i < (int)(sizeof(PrimalFactor) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  double* Ak = A + ((size_t)inst * N + t) * NX * NX;
  double* Bk = Bm + ((size_t)inst * N + t) * NX * NU;
  for (int e = warp; e < NX + NU; e += LINA_WARPS)
    dyn_tangent_id_warp(*md, ws[warp], *pf, xs, xs + NX, e, e < NX ? Ak + e * NX : Bk + (e - NX) * NX);
}

// ---- analytic linearization, one THREAD per column of [A_k | B_k] (h1_lin_dirs.cuh). Threads are packed
//      (knot, direction) -> global thread id, so warps are full regardless of the direction count; the three
//      direction classes are separate instantiations (launches) with their own register budgets.
//      MODE 0: the 26 q columns, 1: the 25 v columns, 2: the 19 control columns. ----
constexpr int LIND_THREADS = 128;
template <int MODE, bool H1TREE>   // H1TREE: the model has H1's dof tree (DynModel::seq_ok) -> static-index solve
__global__ void __launch_bounds__(LIND_THREADS)
k_linearize_dirs(const DynModel* gmd, long nknots, int N, const int* __restrict__ active,
                 const double* __restrict__ xbar, const double* __restrict__ ubar,
                 const PrimalFactor* __restrict__ pf_g, double* __restrict__ A, double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DynModel* md;
  stage_model(smem, gmd, &md);
  constexpr int ND = MODE == 0 ? NQ : (MODE == 1 ? NV : NU);
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long knot = g / ND;
  const int dir = (int)(g - knot * ND);
  if (knot >= nknots) return;
  const long inst = knot / N;
  const int t = (int)(knot - inst * N);
  if (active && !active[inst]) return;
  const double* x = xbar + ((size_t)inst * (N + 1) + t) * NX;
  const PrimalFactor* pf = pf_g + knot;
  double tv[NV];
  int seed;
  if (MODE == 0) { seed = dir; id_tangent_seq<Dual, Dual>(*md, x, pf->a, seed, tv); }
  else if (MODE == 1) { seed = NQ + dir; id_tangent_seq<double, Dual>(*md, x, pf->a, seed, tv); }
  else {
    seed = NX + dir;
    const double uj = ubar[(size_t)knot * NU + dir];
    for (int j = 0; j < NV; ++j) tv[j] = 0.0;
    tv[6 + dir] = (uj < md->ctrl_lo[dir] || uj > md->ctrl_hi[dir]) ? 0.0 : 1.0;   // clamped torque: no sensitivity
  }
  if (H1TREE) tangent_solve_h1(&pf->Lm[0][0], pf->D, tv);
  else tangent_solve_seq(*md, &pf->Lm[0][0], pf->D, tv);
  double* col = (MODE == 2) ? Bm + (size_t)knot * NX * NU + (size_t)dir * NX : A + (size_t)knot * NX * NX + (size_t)seed * NX;
  integrate_tangent_seq(*md, x, pf->a, seed, tv, col);
}

// ---- analytic linearization, one thread per column, DIRECTION-UNIFORM WARPS: a CTA owns LINC_KNOTS = 32
//      consecutive knots and stages their Mhat factors (with 1 / D), states, controls and the Jacobians of the
//      quaternion update in shared memory once (coalesced); each of its warps then repeatedly takes the next work
//      item of the class (shared counter, costliest first) and differentiates the 32 knots along it, one knot per
//      lane. Every lane of a warp therefore runs the same walk — in particular the sparse subtree walks of the joint
//      directions (id_tangent_sub) — and reads its knot's factor from shared memory at an odd stride (conflict
//      free). Columns are staged per warp and written to A_k / B_k as contiguous 408-byte runs.
//      CLS 0: hinge angles (19 columns, subtree walks on dual kinematics)
//          1: base angular velocity (3, every body) + hinge rates (19, subtree walks), plain kinematics
//          2: the light class — base rotations (3 items: rigid-direction tangents, h1_lin_dirs.cuh; kept in shared
//             memory), z and base linear velocity (4 columns, contact only), controls (19, triangular solves
//             only), x / y (unit vectors), and last the 4 raw-quaternion columns = combinations of the three
//             rotation tangents (they wait for the rotation items, which were handed out first).
//      (one launch per class: a single launch over all 70 columns balances the warps better but measured 40 %
//       slower — the warps of a CTA then run five different code paths and thrash the instruction cache) ----
constexpr int LINC_WARPS = 8, LINC_KNOTS = 32, LINC_THREADS = LINC_WARPS * 32;
constexpr int LINC_LIGHT_WARPS = 7;                               // the light class trades one column tile for the rotation tangents
constexpr int LINC_FS = sizeof(PrimalFactor) / sizeof(double);   // 325 doubles: odd stride
constexpr int LINC_QJ = QJ_DIRS * 4 + 1;                          // 29 doubles per knot: odd stride
static_assert(LINC_FS % 2 == 1 && NX % 2 == 1 && NU % 2 == 1 && NV % 2 == 1, "per-knot strides must be odd (bank-conflict-free lane <-> knot access)");
__host__ __device__ constexpr int linc_nitems(int cls) { return cls == 0 ? NB - 1 : cls == 1 ? 3 + NB - 1 : 3 + 4 + NU + 2 + 4; }
__host__ __device__ constexpr int linc_warps(int cls) { return cls == 2 ? LINC_LIGHT_WARPS : LINC_WARPS; }
__host__ __device__ constexpr size_t linc_smem_doubles(int cls) {
  return (size_t)LINC_KNOTS * (LINC_FS + NX + NU + LINC_QJ) + (size_t)linc_warps(cls) * LINC_KNOTS * NX +
         (cls == 2 ? (size_t)3 * LINC_KNOTS * NV : 0);
}
template <int CLS, bool H1TREE>
__global__ void __launch_bounds__(LINC_THREADS)
k_linearize_cols(const DynModel* gmd, long nknots, int N, const int* __restrict__ active,
                 const int* __restrict__ list, const int* __restrict__ list_count,
                 const double* __restrict__ xbar, const double* __restrict__ ubar,
                 const PrimalFactor* __restrict__ pf_g, double* __restrict__ A, double* __restrict__ Bm) {
  extern __shared__ __align__(16) unsigned char smem[];
  // `list` (optional): compact list of the active instances built on the device by k_solve_state; slot s of the
  // grid then works on knot (s % N) of instance list[s / N], so every CTA is full however sparse the active set is.
  __shared__ long kid[LINC_KNOTS];                    // global knot id (instance * N + t) of each slot, -1: nothing to do
  __shared__ int next_item, rot_done;
  constexpr int NW = linc_warps(CLS);
  const long knot0 = (long)blockIdx.x * LINC_KNOTS;
  if (list) nknots = min(nknots, (long)(*list_count) * N);
  if (knot0 >= nknots) return;
  const int nk = (int)min((long)LINC_KNOTS, nknots - knot0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < LINC_KNOTS) {
    long id = -1;
    if (tid < nk) {
      const long sl = knot0 + tid;
      if (list) id = (long)list[sl / N] * N + sl % N;
      else if (!active || active[sl / N]) id = sl;
    }
    kid[tid] = id;
  }
  if (tid == 0) { next_item = 0; rot_done = 0; }
  const DynModel* md;
  unsigned char* p = stage_model(smem, gmd, &md);    // (has a __syncthreads)
  bool any = false;
  for (int k = 0; k < nk; ++k) any |= kid[k] >= 0;
  if (!any) return;                                   // all 32 knots belong to finished instances
  double* fac = reinterpret_cast<double*>(p);        // [LINC_KNOTS][LINC_FS]
  double* xs = fac + LINC_KNOTS * LINC_FS;           // [LINC_KNOTS][NX]
  double* us = xs + LINC_KNOTS * NX;                 // [LINC_KNOTS][NU]
  double* qj = us + LINC_KNOTS * NU;                 // [LINC_KNOTS][LINC_QJ]
  double* tile = qj + LINC_KNOTS * LINC_QJ + warp * LINC_KNOTS * NX;   // this warp's [LINC_KNOTS][NX]
  double* trot = qj + LINC_KNOTS * LINC_QJ + NW * LINC_KNOTS * NX;     // (CLS 2) [3][LINC_KNOTS][NV]
  {
    constexpr int DOFF = NV * MAXSLOT;               // PrimalFactor::D
    for (int i = tid; i < nk * LINC_FS; i += NW * 32) {
      const int k = i / LINC_FS, j = i - k * LINC_FS;
      const long id = kid[k];
      if (id < 0) continue;
      const double v = reinterpret_cast<const double*>(pf_g + id)[j];
      fac[i] = (j >= DOFF && j < DOFF + NV) ? 1.0 / v : v;
    }
    for (int i = tid; i < nk * NX; i += NW * 32) {
      const int k = i / NX, j = i - k * NX;
      const long id = kid[k];
      if (id < 0) continue;
      const long inst = id / N;
      xs[i] = xbar[((size_t)inst * (N + 1) + (id - inst * N)) * NX + j];
    }
    for (int i = tid; i < nk * NU; i += NW * 32) {
      const int k = i / NU, j = i - k * NU;
      const long id = kid[k];
      if (id >= 0) us[i] = ubar[(size_t)id * NU + j];
    }
  }
  __syncthreads();
  if (tid < LINC_KNOTS * QJ_DIRS) {   // Jacobians of the quaternion update: one (knot, direction) per thread
    const int k = tid / QJ_DIRS, d = tid - k * QJ_DIRS;
    if (k < nk && kid[k] >= 0)
      quat_step_jac_dir(*md, xs + k * NX, reinterpret_cast<const PrimalFactor*>(fac + k * LINC_FS)->a, d, qj + k * LINC_QJ + 4 * d);
  }
  __syncthreads();
  const bool ok = lane < nk && kid[lane] >= 0;
  const double* x = xs + lane * NX;
  const PrimalFactor* pf = reinterpret_cast<const PrimalFactor*>(fac + lane * LINC_FS);   // (D holds reciprocals)
  double* col = tile + lane * NX;
  while (true) {
    int it = 0;
    if (lane == 0) it = atomicAdd(&next_item, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= linc_nitems(CLS)) break;
    int seed = -1;                 // output column (0..50 of A, 51..69 -> B); -1: a rotation item (no column)
    if (CLS == 0) seed = 6 + md->dir_order[it];                          // hinges by decreasing subtree size
    else if (CLS == 1) seed = it < 3 ? NQ + 3 + it : NQ + 5 + md->dir_order[it - 3];
    else {
      if (it < 3) seed = -1;
      else if (it == 3) seed = 2;
      else if (it < 7) seed = NQ + it - 4;
      else if (it < 7 + NU) seed = NX + it - 7;
      else if (it < 9 + NU) seed = it - 7 - NU;
      else seed = 3 + it - 9 - NU;
    }
    if (CLS == 2 && seed >= 3 && seed < 7) {   // quaternion column: needs the three rotation tangents
      if (lane == 0) while (atomicAdd(&rot_done, 0) < 3) { }
      __syncwarp();
      __threadfence_block();
    }
    if (ok) {
      if (CLS == 2 && seed >= 0 && seed < 2) {
        for (int j = 0; j < NX; ++j) col[j] = (j == seed) ? 1.0 : 0.0;   // f_D is translation invariant in x and y
      } else if (CLS == 2 && seed < 0) {
        double tv[NV];
        id_tangent_rot(*md, x, pf->a, it, tv);
        double* dst = trot + ((size_t)it * LINC_KNOTS + lane) * NV;
        for (int j = 0; j < NV; ++j) dst[j] = tv[j];
      } else {
        double tv[NV];
        if (CLS == 0) id_tangent_sub<Dual, Dual>(*md, x, pf->a, seed, seed - 6, tv);
        else if (CLS == 1) {
          if (it < 3) id_tangent_seq<double, Dual>(*md, x, pf->a, seed, tv);
          else id_tangent_sub<double, Dual>(*md, x, pf->a, seed, seed - NQ - 5, tv);
        } else if (seed >= NX) {
          const int j = seed - NX;
          const double uj = us[lane * NU + j];
          for (int k = 0; k < NV; ++k) tv[k] = 0.0;
          tv[6 + j] = (uj < md->ctrl_lo[j] || uj > md->ctrl_hi[j]) ? 0.0 : 1.0;   // clamped torque: no sensitivity
        } else if (seed >= 3 && seed < 7) {
          double G[3][4];
          quat_rot_map(x + 3, G);
          const double g0 = G[0][seed - 3], g1 = G[1][seed - 3], g2 = G[2][seed - 3];
          const double* t0 = trot + (size_t)lane * NV;
          for (int j = 0; j < NV; ++j)
            tv[j] = g0 * t0[j] + g1 * t0[LINC_KNOTS * NV + j] + g2 * t0[2 * LINC_KNOTS * NV + j];
        } else {
          id_tangent_rigid(*md, x, pf->a, seed, tv);
        }
        if (H1TREE) tangent_solve_h1<true>(&pf->Lm[0][0], pf->D, tv);
        else tangent_solve_seq<true>(*md, &pf->Lm[0][0], pf->D, tv);
        integrate_tangent_pre(*md, seed, tv, qj + lane * LINC_QJ, col);
      }
    }
    __syncwarp();
    if (CLS == 2 && seed < 0) {    // publish the rotation tangent
      __threadfence_block();
      if (lane == 0) atomicAdd(&rot_done, 1);
      continue;
    }
    // the 32 columns of this item: contiguous 51-double runs
    {
      double* dst0 = (seed >= NX) ? Bm + (size_t)(seed - NX) * NX : A + (size_t)seed * NX;
      const size_t kstride = (seed >= NX) ? (size_t)NX * NU : (size_t)NX * NX;
      int k = 0, j = lane;
#pragma unroll 1
      for (int e = lane; e < nk * NX; e += 32) {
        const long id = kid[k];
        if (id >= 0) dst0[(size_t)id * kstride + j] = tile[e];
        j += 32;
        if (j >= NX) { j -= NX; ++k; }
      }
    }
    __syncwarp();
  }
}

}  // namespace h1
