// Shared definitions for the sm_100a kernels of the H1 iLQR hot path.
#pragma once
#include "../../include/h1ilqr.h"

#if defined(__CUDACC__)
#define H1_DEV __device__ __forceinline__
#define H1_HD __host__ __device__ __forceinline__
#else
// Lane-emulation build (tests/emul): the warp-cooperative phase functions are plain C++ and are run
// lane-by-lane between the points where the kernel has a __syncwarp().
#define H1_DEV inline
#define H1_HD inline
#include <cmath>
namespace h1 { using std::isfinite; }
#endif

namespace h1 {

#if defined(__CUDACC__)
// D(8x8) += A(8x4) B(4x8) on the fp64 tensor core (SASS DMMA). Lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

constexpr int NB = H1_NB, NQ = H1_NQ, NV = H1_NV, NX = H1_NX, NU = H1_NU;

// Device-side weights: the C-ABI struct (diagonals of Q / R / Qf + scalar task weights) followed by an optional pointer to
// the OFF-DIAGONAL parts of symmetric full Q, R, Qf (iLQR multiplies full matrices, ilqr.cpp:145-150, 372-373, 441):
// [Qoff 51x51 | Roff 19x19 | Qfoff 51x51], column-major, zero diagonals; nullptr = the matrices are diagonal (what
// Config::buildCostMatrices builds). Every kernel receives `const H1Weights*` that points at the first member of one of
// these, so the cost functions reach the pointer without a change of their signatures.
struct DevWeights {
  H1Weights w;
  const double* qoff;
};
// Per-knot strides of A_k, B_k, lxx_k in device memory: the dense column-major matrices (leading dimension NX) followed by ONE
// pad double, so that every matrix starts on a 16-byte boundary and spans a multiple of 16 bytes — the Riccati kernel
// fetches them with one bulk-copy instruction each (cp.async.bulk needs both). The pad is never written (zero).
constexpr int A_STRIDE = NX * NX + 1, B_STRIDE = NX * NU + 1, LXX_STRIDE = NX * NX + 1;
static_assert((A_STRIDE * 8) % 16 == 0 && (B_STRIDE * 8) % 16 == 0, "bulk copies need 16-byte multiples");
constexpr int QOFF_R = NX * NX, QOFF_QF = NX * NX + NU * NU, QOFF_SIZE = 2 * NX * NX + NU * NU;
H1_HD const double* weights_offdiag(const H1Weights& wt) {
#if defined(__CUDACC__)
  return reinterpret_cast<const DevWeights&>(wt).qoff;
#else
  (void)wt;
  return nullptr;   // (emulation builds hand over a bare H1Weights)
#endif
}
constexpr int MAXSLOT = 11;  // base 6 + longest hinge chain 5
constexpr int NCPT = H1_NFOOT * H1_NCP;

// Read-only kinematic tree + parameters of the dynamics model, built on the host from H1Model
// (model_tables.cpp) and staged into shared memory by every kernel that evaluates f_D.
struct DynModel {
  double pos[NB][3];
  double rfix[NB][9];
  double ipos[NB][3];
  double inertia[NB][6];
  double mass[NB];
  double armature[NV], damping[NV];
  double ctrl_lo[NU], ctrl_hi[NU];
  double jnt_lo[NU], jnt_hi[NU];
  double foot_pts[NCPT][3];
  double gravity[3];
  double h, kn, bn, bt, eps, total_mass, inv_total_mass;
  int parent[NB], axis[NB], has_rfix[NB], depth[NB];
  int anc_body[NB][6];       // anc_body[b][d]: ancestor of b at depth d (d = depth[b] -> b itself)
  int nlist[NV];             // #slots of dof j: its ancestor dofs root->self, self included
  int alist[NV][MAXSLOT];    // dof ids of those slots
  int level[NV];             // depth of dof j in the dof tree (0..10)
  int dof_sub_end[NV];       // last dof of the dof-subtree rooted at j (dof subtrees are contiguous)
  int chain_end[NB];         // last body of b's subtree (bodies are in DFS order)
  int foot_body[H1_NFOOT];
  int foot_dof[H1_NFOOT];    // ankle dof of each foot
  int cp_lo[NV], cp_hi[NV];  // contact points [cp_lo, cp_hi) move with dof j
  int base_child_slot[NB];   // index of body b among the base's children (only valid when parent[b] == 0)
  int n_base_children;
  int pad_;
  int nchild[NB];            // number of child bodies (branch bodies keep their state in the sequential walks)
  int dir_order[NB - 1];     // hinged bodies by decreasing subtree size (costliest linearization directions first)
  // work order of the merged tangent kernel (h1_lin_finish.cuh): all 49 direction items of a knot group, class by class
  // (class << 5 | item of that class), costliest first within a class (model_tables.cpp)
  unsigned char tan_order[52];
  int n_tan_items;
  int seq_ok;                // 1: the tree has H1's chain structure the thread-sequential f_D (h1_dyn_seq.cuh) is specialised for
};

// Compile-time copy of H1's dof tree (valid when DynModel::seq_ok): slot count of dof k and the dof id in slot s.
// Base dofs 0-5 form a chain; dofs 6-10 / 11-15 = legs, 16 = torso, 17-20 / 21-24 = arms below the torso.
constexpr int h1_nlist(int k) { return k < 6 ? k + 1 : (k < 16 ? 7 + (k - 6) % 5 : (k == 16 ? 7 : 8 + (k - 17) % 4)); }
constexpr int h1_anc(int k, int s) {
  return s < 6 ? s : (k < 16 ? 6 + 5 * ((k - 6) / 5) + (s - 6) : ((k == 16 || s == 6) ? 16 : 17 + 4 * ((k - 17) / 4) + (s - 7)));
}

// Same tree for the cost (URDF / Pinocchio-semantics) model; only what the cost kernel reads.
struct CostModel {
  double pos[NB][3];
  double rfix[NB][9];
  double ipos[NB][3];
  double wmass[NB];          // mass[b] / total_mass
  int parent[NB], axis[NB], has_rfix[NB], depth[NB];
  int anc_body[NB][6];
  int chain_end[NB];
  int foot_body[H1_NFOOT];
  // ordered joint pairs (k <= l, bodies 1..NB-1) of the Hessian contraction tables: the first n_anc_pairs are the pairs
  // with k an ancestor-or-self of l (59 for H1) — only these have non-zero entries — so a warp computes them with full
  // lanes instead of scanning all 19 x 19 combinations
  int n_anc_pairs;
  unsigned char pair_k[(NB - 1) * NB / 2 + 2], pair_l[(NB - 1) * NB / 2 + 2];
};
constexpr int CQ_NPAIRS = (NB - 1) * NB / 2;

}  // namespace h1
