// C ABI of the B200-native H1 iLQR solver core (include/h1ilqr.h): handle, device-resident buffers, kernel
// launch sequences. Host language above this file is C++ (host/) or Python ctypes (tests, bench).
// There is no CPU fallback anywhere in this library: every entry point runs CUDA kernels on the handle's
// device or returns H1ILQR_ECUDA.
#include "h1_kernels_solve.cuh"
#include "h1_kernels_seq.cuh"
#include "h1_kernels_quad.cuh"
#include "h1_lin_finish.cuh"
#include "h1_riccati.cuh"
#include "model_tables.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

using namespace h1;

static thread_local std::string g_err;
static int set_err(int code, const char* what, cudaError_t e = cudaSuccess) {
  g_err = what;
  if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
  return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(H1ILQR_ECUDA, #call, e_); } while (0)

struct H1Ilqr {
  int B = 0, N = 0, device = 0;
  cudaStream_t stream = nullptr;        // every launch goes to `stream` (enqueue_solve swaps in `stream2` for the second attempts)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  H1SolverOptions opt;
  DynModel* d_dyn = nullptr; CostModel* d_cost = nullptr; H1Weights* d_w = nullptr; H1SolverOptions* d_opt = nullptr;
  DevWeights* d_dw = nullptr;   // d_w points at its first member; qoff (below) = off-diagonal parts of full Q / R / Qf or nullptr
  double* d_qoff = nullptr; bool qoff_on = false;
  double *xbar = nullptr, *ubar = nullptr, *K = nullptr, *kff = nullptr, *A = nullptr, *Bm = nullptr;
  double *lx = nullptr, *lu = nullptr, *lxx = nullptr, *luu = nullptr, *xnew = nullptr, *unew = nullptr;
  double *x0 = nullptr, *u_init = nullptr, *u_apply = nullptr, *prev_xbar = nullptr, *prev_ubar = nullptr;
  double *x_ref = nullptr, *u_ref = nullptr, *com_ref = nullptr, *ee_ref = nullptr, *com_vel_ref = nullptr;
  int* stance = nullptr; int ref_shared = 1;
  double *lambda = nullptr, *cost = nullptr, *prev_cost = nullptr, *nominal_cost = nullptr, *ls_cost = nullptr;
  int *active = nullptr, *second = nullptr, *iters = nullptr, *status = nullptr, *ls_ok = nullptr, *ls_alpha = nullptr;
  int *act_list = nullptr, *sec_list = nullptr, *list_count = nullptr;
  int *early = nullptr, *late = nullptr, *early_list = nullptr, *late_list = nullptr;
  int *has_prev = nullptr, *warm_mask = nullptr, *cold_mask = nullptr, *warm_in = nullptr;
  double* cost_trace = nullptr; int* alpha_trace = nullptr;
  PrimalFactor* pf = nullptr;   // [B][N] factorisation of Mhat at every knot of the nominal trajectory
  double* scratch = nullptr; size_t scratch_bytes = 0;   // device staging for n-state queries
  void* pin = nullptr; size_t pin_bytes = 0;              // pinned host staging
  std::vector<void*> allocs;
  // timing
  bool timing = false;
  H1StageTimes times;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int launches = 0;
  size_t smem_lina = 0;
  int policy = H1ILQR_KERNELS_AUTO;
  int seq_min_batch = 768;  // AUTO: batch at or above which rollouts / line searches run one THREAD per f_D evaluation (h1_dyn_seq.cuh);
                            // measured break-even with the warp-per-evaluation kernels: between 512 and 1024 instances
  size_t smem_seq = 0, smem_seq_ls = 0, smem_lint = 0, smem_linf = 0, smem_q4 = 0;
  int q4_warps = 8;         // instances per CTA of k_line_search_quad
  bool roll_quad = true;    // batched nominal rollout: quad-cooperative kernel; false = one thread per instance
  size_t smem_rq = 0;
  bool ls_quad = true;      // batched line search: quad-cooperative kernel (h1_kernels_quad.cuh); false = thread-sequential one
  bool seq_ok = false;      // the model has the chain structure the thread-sequential f_D is specialised for
  long lin_cols_min_knots = 148 * 32;   // AUTO: B*N at or above which the direction-uniform linearization (32 knots per CTA) fills the GPU;
                                        // below it the knot-major thread-per-column kernel (k_linearize_dirs) is the faster one
  size_t smem_dyn4 = 0, smem_lin = 0, smem_cq = 0, smem_ls = 0, smem_ric = 0;
  // Riccati: [A|B] in h->A / h->Bm were written by an analytic linearization kernel (position rows = unit entry + dt * velocity
  // rows, h1_riccati.cuh) -> the contractions run over the 29 reduced rows. Cleared by the forward-difference kernel and by
  // h1ilqr_set_linearization; H1_RIC_DENSE=1 keeps the dense kernel (A/B measurements)
  bool ab_struct = false, ric_struct_ok = true;
  bool lint_merged = true;   // tangent directions of all three classes in one launch (h1_lin_finish.cuh)
  double dt = 0.0;
  // device-resident closed loop: full reference tables, per-instance time index, step counter, per-step logs
  double *tab_x = nullptr, *tab_com = nullptr, *tab_ee = nullptr, *tab_cv = nullptr;
  int* tab_contact = nullptr;
  int tab_rows = 0, tab_contact_rows = 0, tab_schedule_offset = 0;
  int *t_idx = nullptr, *step_ctr = nullptr;
  double* plant_next = nullptr;
  std::vector<void*> table_allocs;
  // CUDA graphs of one step: [0] resident warm, [1] resident cold, [2] closed loop (without logs), [3] closed loop (with logs)
  cudaGraphExec_t graph[4] = {nullptr, nullptr, nullptr, nullptr};
  int graph_launches[4] = {0, 0, 0, 0};
  double *log_cost = nullptr, *log_u = nullptr; int* log_iters = nullptr; int log_capacity = 0;
};

template <class T> static cudaError_t dalloc(H1Ilqr* h, T** p, size_t n) {
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e == cudaSuccess) { h->allocs.push_back(*p); e = cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream); }
  return e;
}

static void invalidate_graphs(H1Ilqr* h) {   // pointers / flags baked into a captured step changed
  for (auto& g : h->graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}
static RefTable ref_table(const H1Ilqr* h) {
  RefTable r;
  r.x_ref = h->x_ref; r.u_ref = h->u_ref; r.com_ref = h->com_ref; r.ee_ref = h->ee_ref;
  r.com_vel_ref = h->com_vel_ref; r.stance = h->stance; r.shared = h->ref_shared; r.N = h->N;
  return r;
}

extern "C" {

void h1ilqr_default_options(H1SolverOptions* o) {
  o->max_iterations = 10; o->tolerance = 1e-4; o->reg_init = 1e-6; o->reg_min = 1e-6; o->reg_max = 1e-3;
  o->accept_margin = 1e-6; o->fd_eps = 1e-5; o->divergence_cost = 1e6; o->linearization = H1ILQR_LIN_ANALYTIC;
  const double a[H1ILQR_NALPHA] = {1.0, 0.8, 0.6, 0.4, 0.2, 0.1, 0.05, 0.01};
  std::memcpy(o->alphas, a, sizeof(a));
}

const char* h1ilqr_last_error(void) { return g_err.c_str(); }
int h1ilqr_batch(const H1Ilqr* h) { return h ? h->B : 0; }
int h1ilqr_horizon(const H1Ilqr* h) { return h ? h->N : 0; }
void* h1ilqr_stream(H1Ilqr* h) { return h ? (void*)h->stream : nullptr; }

void h1ilqr_destroy(H1Ilqr* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& g : h->graph) if (g) cudaGraphExecDestroy(g);
  for (void* p : h->table_allocs) cudaFree(p);
  for (void* p : h->allocs) cudaFree(p);
  if (h->pin) cudaFreeHost(h->pin);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int h1ilqr_create(const H1Model* dyn_model, const H1Model* cost_model, const H1SolverOptions* opt, int batch, int N,
                  int device, H1Ilqr** out) {
  if (!out || batch < 1 || N < 2) return set_err(H1ILQR_EARG, "h1ilqr_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return set_err(H1ILQR_ECUDA, "no CUDA device (this library has no CPU fallback)", e);
  if (device < 0 || device >= ndev) return set_err(H1ILQR_EARG, "h1ilqr_create: bad device index");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return set_err(H1ILQR_ECUDA, "device is not sm_100 class (kernels are built for sm_100a only)");
  H1Ilqr* h = new H1Ilqr;
  h->B = batch; h->N = N; h->device = device;
  if (opt) h->opt = *opt; else h1ilqr_default_options(&h->opt);
  if (h->opt.max_iterations < 1 || h->opt.max_iterations > H1ILQR_MAX_ITERS) { delete h; return set_err(H1ILQR_EARG, "max_iterations out of range"); }
  DynModel dm; CostModel cm;
  if (!build_dyn_model(dyn_model ? *dyn_model : *h1_default_dynamics_model(), &dm) ||
      !build_cost_model(cost_model ? *cost_model : *h1_default_cost_model(), &cm)) {
    delete h; return set_err(H1ILQR_EARG, "model is not a DFS-ordered H1-like tree");
  }
  h->seq_ok = dm.seq_ok != 0;
  h->dt = dm.h;
  if (const char* e = getenv("H1_RIC_DENSE")) h->ric_struct_ok = atoi(e) == 0;
  if (const char* e = getenv("H1_SEQ_MIN_BATCH")) h->seq_min_batch = atoi(e);   // tuning experiments
#define CUH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(H1ILQR_ECUDA, #call, e_); h1ilqr_destroy(h); return H1ILQR_ECUDA; } } while (0)
  CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUH(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CUH(cudaEventCreate(&h->ev[0])); CUH(cudaEventCreate(&h->ev[1]));
  CUH(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)); CUH(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  const size_t B = batch, N1 = N + 1;
  CUH(dalloc(h, &h->d_dyn, 1)); CUH(dalloc(h, &h->d_cost, 1)); CUH(dalloc(h, &h->d_dw, 1)); CUH(dalloc(h, &h->d_opt, 1));
  h->d_w = &h->d_dw->w;   // (zero-initialised: qoff == nullptr -> diagonal weights)
  CUH(dalloc(h, &h->d_qoff, (size_t)QOFF_SIZE));
  CUH(cudaMemcpyAsync(h->d_dyn, &dm, sizeof(dm), cudaMemcpyHostToDevice, h->stream));
  CUH(cudaMemcpyAsync(h->d_cost, &cm, sizeof(cm), cudaMemcpyHostToDevice, h->stream));
  CUH(cudaMemcpyAsync(h->d_opt, &h->opt, sizeof(h->opt), cudaMemcpyHostToDevice, h->stream));
  CUH(cudaStreamSynchronize(h->stream));  // dm / cm are stack objects
  CUH(dalloc(h, &h->xbar, B * N1 * NX)); CUH(dalloc(h, &h->ubar, B * N * NU));
  CUH(dalloc(h, &h->K, B * N * NU * NX)); CUH(dalloc(h, &h->kff, B * N * NU));
  CUH(dalloc(h, &h->A, B * N * A_STRIDE)); CUH(dalloc(h, &h->Bm, B * N * B_STRIDE));   // padded per-knot strides (h1_common.cuh)
  CUH(dalloc(h, &h->lx, B * N1 * NX)); CUH(dalloc(h, &h->lu, B * N * NU));
  CUH(dalloc(h, &h->lxx, B * N1 * LXX_STRIDE)); CUH(dalloc(h, &h->luu, B * N * NU * NU));
  CUH(dalloc(h, &h->xnew, B * H1ILQR_NALPHA * N1 * NX)); CUH(dalloc(h, &h->unew, B * H1ILQR_NALPHA * N * NU));
  CUH(dalloc(h, &h->x0, B * NX)); CUH(dalloc(h, &h->u_init, B * NU)); CUH(dalloc(h, &h->u_apply, B * NU));
  CUH(dalloc(h, &h->prev_xbar, B * N1 * NX)); CUH(dalloc(h, &h->prev_ubar, B * N * NU));
  CUH(dalloc(h, &h->x_ref, B * N1 * NX)); CUH(dalloc(h, &h->u_ref, B * N * NU)); CUH(dalloc(h, &h->com_ref, B * N1 * 3));
  CUH(dalloc(h, &h->ee_ref, B * N1 * 6)); CUH(dalloc(h, &h->com_vel_ref, B * N1 * 3)); CUH(dalloc(h, &h->stance, B * N1 * 2));
  CUH(dalloc(h, &h->lambda, B)); CUH(dalloc(h, &h->cost, B)); CUH(dalloc(h, &h->prev_cost, B));
  CUH(dalloc(h, &h->nominal_cost, B)); CUH(dalloc(h, &h->ls_cost, B));
  CUH(dalloc(h, &h->active, B)); CUH(dalloc(h, &h->second, B)); CUH(dalloc(h, &h->iters, B)); CUH(dalloc(h, &h->status, B));
  CUH(dalloc(h, &h->ls_ok, B)); CUH(dalloc(h, &h->ls_alpha, B)); CUH(dalloc(h, &h->has_prev, B));
  CUH(dalloc(h, &h->warm_mask, B)); CUH(dalloc(h, &h->act_list, B)); CUH(dalloc(h, &h->sec_list, B)); CUH(dalloc(h, &h->list_count, 4));
  CUH(dalloc(h, &h->early, B)); CUH(dalloc(h, &h->late, B)); CUH(dalloc(h, &h->early_list, B)); CUH(dalloc(h, &h->late_list, B)); CUH(dalloc(h, &h->cold_mask, B)); CUH(dalloc(h, &h->warm_in, B));
  CUH(dalloc(h, &h->pf, B * N));
  CUH(dalloc(h, &h->t_idx, B)); CUH(dalloc(h, &h->step_ctr, 1)); CUH(dalloc(h, &h->plant_next, B * NX));
  CUH(dalloc(h, &h->cost_trace, B * h->opt.max_iterations)); CUH(dalloc(h, &h->alpha_trace, B * h->opt.max_iterations * 2));
  h->scratch_bytes = B * N1 * (NX + NU + NX + NV + 9) * sizeof(double);
  { double* sp = nullptr; CUH(dalloc(h, &sp, h->scratch_bytes / sizeof(double))); h->scratch = sp; }
  h->pin_bytes = B * (NX + 2 * NU + 4) * sizeof(double);
  CUH(cudaMallocHost(&h->pin, h->pin_bytes));
  k_fill_double<<<(batch + 255) / 256, 256, 0, h->stream>>>(batch, h->lambda, h->opt.reg_init);
  k_fill_int<<<(int)((B * N1 * 2 + 255) / 256), 256, 0, h->stream>>>((int)(B * N1 * 2), h->stance, 1);
  // dynamic shared memory sizes
  const size_t mdl = ((sizeof(DynModel) + 15) / 16) * 16, cml = ((sizeof(CostModel) + 15) / 16) * 16;
  h->smem_dyn4 = mdl + 4 * sizeof(DynWarp);
  h->smem_lin = mdl + LIN_WARPS * sizeof(DynWarp) + (LIN_EVALS * NX + NX + NU) * sizeof(double);
  h->smem_lina = mdl + LINA_WARPS * sizeof(TanWarpT<Dual>) + sizeof(PrimalFactor) + (NX + NU) * sizeof(double);
  h->smem_cq = cml + CQ_KNOTS * sizeof(CostWarp);
  h->smem_ls = mdl + H1ILQR_NALPHA * sizeof(DynWarp) + (H1ILQR_NALPHA + H1ILQR_NALPHA * (NX + NU)) * sizeof(double);
  h->smem_ric = sizeof(RiccatiSmem);
  h->smem_seq = mdl;
  h->smem_lint = mdl + LINT_SMEM_DOUBLES * sizeof(double);
  h->smem_linf = LINF_WARPS * sizeof(LinFinishWarp);
  CUH(cudaFuncSetAttribute(k_linearize_tangents<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_lint));
  CUH(cudaFuncSetAttribute(k_linearize_tangents<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_lint));
  CUH(cudaFuncSetAttribute(k_linearize_tangents<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_lint));
  CUH(cudaFuncSetAttribute(k_linearize_tangents<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_lint));
  if (const char* e = getenv("H1_LINT_SPLIT")) h->lint_merged = atoi(e) == 0;   // A/B measurements: one launch per direction class
  CUH(cudaFuncSetAttribute(k_linearize_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_linf));
  CUH(cudaFuncSetAttribute(k_linearize_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_linf));
  if (const char* e = getenv("H1_SEQ_SMEM_PAD")) h->smem_seq += (size_t)atoi(e) * 1024;   // experiment: limits resident CTAs
  h->smem_seq_ls = h->smem_seq + (size_t)SEQ_THREADS * NX * sizeof(double);
  CUH(cudaFuncSetAttribute(k_line_search_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_seq_ls));
  // instances per CTA of the quad line search: 8 (one 215 KB CTA per SM) for large batches; 1 while one-warp CTAs (28 KB,
  // seven per SM) still cover the batch in a single wave, so that a handful of instances spread over as many SMs
  h->q4_warps = batch > 7 * prop.multiProcessorCount ? 8 : 1;
  if (const char* e = getenv("H1_Q4_WARPS")) h->q4_warps = atoi(e) == 1 ? 1 : 8;   // A/B measurements
  h->smem_q4 = mdl + (size_t)h->q4_warps * sizeof(Q4WarpSmem);
  CUH(cudaFuncSetAttribute(k_line_search_quad<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(mdl + 8 * sizeof(Q4WarpSmem))));
  CUH(cudaFuncSetAttribute(k_line_search_quad<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(mdl + 1 * sizeof(Q4WarpSmem))));
  h->smem_rq = mdl + (size_t)RQ_WARPS * sizeof(RQWarpSmem);
  CUH(cudaFuncSetAttribute(k_rollout_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_rq));
  if (const char* e = getenv("H1_ROLL_SEQ")) h->roll_quad = atoi(e) == 0;
  if (const char* e = getenv("H1_LS_SEQ")) h->ls_quad = atoi(e) == 0;   // A/B measurements against the thread-sequential kernel
  CUH(cudaFuncSetAttribute(k_rollout_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_seq));
  CUH(cudaFuncSetAttribute(k_dyn_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_dyn4));
  CUH(cudaFuncSetAttribute(k_dyn_query, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_dyn4));
  CUH(cudaFuncSetAttribute(k_stage_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_dyn4));
  CUH(cudaFuncSetAttribute(k_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_dyn4));
  CUH(cudaFuncSetAttribute(k_primal_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_dyn4));
  CUH(cudaFuncSetAttribute(k_linearize_fd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_lin));
  CUH(cudaFuncSetAttribute(k_linearize_analytic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_lina));
  CUH(cudaFuncSetAttribute(k_cost_quadratics<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_cq));
  CUH(cudaFuncSetAttribute(k_cost_quadratics<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_cq));
  CUH(cudaFuncSetAttribute(k_line_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ls));
  CUH(cudaFuncSetAttribute(k_backward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ric));
  CUH(cudaFuncSetAttribute(k_backward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ric));
  CUH(cudaStreamSynchronize(h->stream));
  CUH(cudaGetLastError());
#undef CUH
  std::memset(&h->times, 0, sizeof(h->times));
  *out = h;
  return H1ILQR_OK;
}

static int guard(H1Ilqr* h) {
  if (!h) return set_err(H1ILQR_EARG, "null handle");
  CU(cudaSetDevice(h->device));
  return 0;
}
#define GUARD(h) do { int g_ = guard(h); if (g_) return g_; } while (0)
#define H2D(dst, src, n) CU(cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, h->stream))
#define D2H(dst, src, n) CU(cudaMemcpyAsync(dst, src, (n), cudaMemcpyDeviceToHost, h->stream))
#define SYNC() CU(cudaStreamSynchronize(h->stream))
#define LAUNCHED() do { ++h->launches; } while (0)

int h1ilqr_set_weights(H1Ilqr* h, const H1Weights* w) {
  GUARD(h);
  if (!w) return set_err(H1ILQR_EARG, "null weights");
  H2D(h->d_w, w, sizeof(*w));
  SYNC();
  return 0;
}

// Full symmetric Q / R / Qf (RobotUtils::setCostWeights keeps whole matrices and iLQR multiplies them, ilqr.cpp:145-150,
// 372-373, 441). The diagonals go to the H1Weights arrays, the off-diagonal parts to a device table that every cost kernel
// adds when present. Column-major 51x51 / 19x19 / 51x51; NULL for all three = back to diagonal weights.
int h1ilqr_set_weight_matrices(H1Ilqr* h, const double* Q, const double* R, const double* Qf) {
  GUARD(h);
  if (!Q && !R && !Qf) {
    const double* none = nullptr;
    H2D(&h->d_dw->qoff, &none, sizeof(none));
    SYNC();
    if (h->qoff_on) invalidate_graphs(h);
    h->qoff_on = false;
    return 0;
  }
  if (!Q || !R || !Qf) return set_err(H1ILQR_EARG, "h1ilqr_set_weight_matrices: give all three matrices or none");
  auto symmetric = [](const double* M, int n) {
    for (int j = 0; j < n; ++j) for (int i = 0; i < j; ++i) if (M[j * n + i] != M[i * n + j]) return false;
    return true;
  };
  if (!symmetric(Q, NX) || !symmetric(R, NU) || !symmetric(Qf, NX))
    return set_err(H1ILQR_EARG, "h1ilqr_set_weight_matrices: Q, R, Qf must be symmetric");
  H1Weights w;
  CU(cudaMemcpyAsync(&w, h->d_w, sizeof(w), cudaMemcpyDeviceToHost, h->stream));
  SYNC();
  std::vector<double> off(QOFF_SIZE);
  for (int j = 0; j < NX; ++j) for (int i = 0; i < NX; ++i) {
    off[j * NX + i] = (i == j) ? 0.0 : Q[j * NX + i];
    off[QOFF_QF + j * NX + i] = (i == j) ? 0.0 : Qf[j * NX + i];
  }
  for (int j = 0; j < NU; ++j) for (int i = 0; i < NU; ++i) off[QOFF_R + j * NU + i] = (i == j) ? 0.0 : R[j * NU + i];
  for (int i = 0; i < NX; ++i) { w.Qdiag[i] = Q[i * NX + i]; w.Qfdiag[i] = Qf[i * NX + i]; }
  for (int i = 0; i < NU; ++i) w.Rdiag[i] = R[i * NU + i];
  bool any = false;
  for (double v : off) any |= (v != 0.0);
  H2D(h->d_w, &w, sizeof(w));
  H2D(h->d_qoff, off.data(), off.size() * sizeof(double));
  const double* ptr = any ? h->d_qoff : nullptr;
  H2D(&h->d_dw->qoff, &ptr, sizeof(ptr));
  SYNC();
  if (h->qoff_on != any) invalidate_graphs(h);   // the table pointer is a kernel argument of the cost-quadratics launch
  h->qoff_on = any;
  return 0;
}

int h1ilqr_set_reference_window(H1Ilqr* h, const double* x_ref, const double* u_ref, const double* com_ref,
                                const double* ee_ref, const int* stance, const double* com_vel_ref, int shared) {
  GUARD(h);
  if (!x_ref || !u_ref || !com_ref || !ee_ref || !stance) return set_err(H1ILQR_EARG, "null reference array");
  const size_t n = shared ? 1 : h->B, N1 = h->N + 1, N = h->N;
  if (h->ref_shared != (shared ? 1 : 0)) invalidate_graphs(h);   // the flag is a kernel argument of a captured step
  h->ref_shared = shared ? 1 : 0;
  H2D(h->x_ref, x_ref, n * N1 * NX * sizeof(double)); H2D(h->u_ref, u_ref, n * N * NU * sizeof(double));
  H2D(h->com_ref, com_ref, n * N1 * 3 * sizeof(double)); H2D(h->ee_ref, ee_ref, n * N1 * 6 * sizeof(double));
  H2D(h->stance, stance, n * N1 * 2 * sizeof(int));
  if (com_vel_ref) H2D(h->com_vel_ref, com_vel_ref, n * N1 * 3 * sizeof(double));
  else CU(cudaMemsetAsync(h->com_vel_ref, 0, n * N1 * 3 * sizeof(double), h->stream));
  SYNC();
  return 0;
}

// ---------------- stage launchers (no host sync) ----------------
// compact instance list (and its device-side counter) that k_solve_state maintains for one of the solve masks
static const int* list_for(const H1Ilqr* h, const int* mask, const int** count) {
  const int* list = nullptr; int slot = 0;
  if (mask && mask == h->active) { list = h->act_list; slot = 0; }
  else if (mask && mask == h->second) { list = h->sec_list; slot = 1; }
  else if (mask && mask == h->early) { list = h->early_list; slot = 2; }
  else if (mask && mask == h->late) { list = h->late_list; slot = 3; }
  *count = list ? h->list_count + slot : nullptr;
  return list;
}
static bool use_batched(const H1Ilqr* h, long units, long auto_min_units) {
  if (h->policy == H1ILQR_KERNELS_COOPERATIVE) return false;
  if (h->policy == H1ILQR_KERNELS_BATCHED) return true;
  return units >= auto_min_units;
}
static void launch_factors(H1Ilqr* h, const int* mask);
static void launch_rollout(H1Ilqr* h, const int* mask, const double* x0_dev, int t_begin, double* cost_out,
                           bool keep_factors = false) {
  if (h->seq_ok && h->roll_quad && use_batched(h, h->B, h->seq_min_batch)) {   // four lanes per instance; factors knot-parallel
    const int per_cta = RQ_WARPS * 8;
    k_rollout_quad<<<(h->B + per_cta - 1) / per_cta, RQ_WARPS * 32, h->smem_rq, h->stream>>>(
        h->d_dyn, h->d_w, ref_table(h), h->B, h->N, t_begin, mask, x0_dev, h->xbar, h->ubar, cost_out);
    LAUNCHED();
    if (keep_factors) launch_factors(h, mask);
    return;
  }
  if (h->seq_ok && use_batched(h, h->B, h->seq_min_batch)) {   // one thread per instance
    k_rollout_seq<<<(h->B + SEQ_ROLL_THREADS - 1) / SEQ_ROLL_THREADS, SEQ_ROLL_THREADS, h->smem_seq, h->stream>>>(
        h->d_dyn, h->d_w, ref_table(h), h->B, h->N, t_begin, mask, x0_dev, h->xbar, h->ubar, cost_out,
        keep_factors ? h->pf : nullptr);
    LAUNCHED();
    return;
  }
  const int wpb = 4, blocks = (h->B + wpb - 1) / wpb;
  k_rollout<<<blocks, wpb * 32, h->smem_dyn4, h->stream>>>(h->d_dyn, h->d_w, ref_table(h), h->B, h->N, t_begin, mask,
                                                          x0_dev, h->xbar, h->ubar, cost_out,
                                                          keep_factors ? h->pf : nullptr);
  LAUNCHED();
}
// Mhat factors at every knot of the current trajectory (knot-parallel)
static void launch_factors(H1Ilqr* h, const int* mask) {
  const long knots = (long)h->B * h->N;
  if (h->seq_ok && use_batched(h, h->B, h->seq_min_batch)) {
    const int* cnt; const int* list = list_for(h, mask, &cnt);
    k_primal_factor_seq<<<(unsigned)((knots + SEQ_ROLL_THREADS - 1) / SEQ_ROLL_THREADS), SEQ_ROLL_THREADS, h->smem_seq, h->stream>>>(
        h->d_dyn, h->B, h->N, mask, list, cnt, h->xbar, h->ubar, h->pf);
  } else {
    k_primal_factor<<<(int)((knots + 3) / 4), 128, h->smem_dyn4, h->stream>>>(h->d_dyn, h->B, h->N, mask, h->xbar, h->ubar, h->pf);
  }
  LAUNCHED();
}
// `factors_ready`: the factorisations of Mhat of the current trajectory are already in h->pf
static void launch_linearize(H1Ilqr* h, const int* mask, bool factors_ready = false) {
  h->ab_struct = h->opt.linearization != H1ILQR_LIN_FD;
  if (h->opt.linearization != H1ILQR_LIN_FD) {
    if (!factors_ready) launch_factors(h, mask);
    const long knots = (long)h->B * h->N;
    if (use_batched(h, knots, h->lin_cols_min_knots)) {  // one thread per column, direction-uniform warps
      const unsigned kb = (unsigned)((knots + LINT_KNOTS - 1) / LINT_KNOTS);
      const int* cnt; const int* list = list_for(h, mask, &cnt);   // compact instance list (k_solve_state)
      // tangents parked in A_k, then one dense contraction per knot (h1_lin_finish.cuh)
#define LINT_LAUNCH(CLS) \
  k_linearize_tangents<CLS><<<kb, LINT_THREADS, h->smem_lint, h->stream>>>(h->d_dyn, knots, h->N, mask, list, cnt, h->xbar, h->pf, h->A)
      if (h->lint_merged) { LINT_LAUNCH(3); h->launches -= 2; }
      else { LINT_LAUNCH(0); LINT_LAUNCH(1); LINT_LAUNCH(2); }
#undef LINT_LAUNCH
      const unsigned fb = (unsigned)((knots + LINF_WARPS - 1) / LINF_WARPS);
      if (h->seq_ok) k_linearize_finish<true><<<fb, LINF_THREADS, h->smem_linf, h->stream>>>(h->d_dyn, knots, h->N, mask, list, cnt, h->xbar, h->ubar, h->pf, h->A, h->Bm);
      else k_linearize_finish<false><<<fb, LINF_THREADS, h->smem_linf, h->stream>>>(h->d_dyn, knots, h->N, mask, list, cnt, h->xbar, h->ubar, h->pf, h->A, h->Bm);
      h->launches += 4;
      return;
    }
    if (h->policy == H1ILQR_KERNELS_AUTO) {   // small batches: one thread per column, knot-major (the factor is shared by a warp)
      const size_t sm = ((sizeof(DynModel) + 15) / 16) * 16;
      auto blocks = [&](int nd) { return (unsigned)((knots * nd + LIND_THREADS - 1) / LIND_THREADS); };
#define LIND_LAUNCH(MODE, ND, TREE) \
  k_linearize_dirs<MODE, TREE><<<blocks(ND), LIND_THREADS, sm, h->stream>>>(h->d_dyn, knots, h->N, mask, h->xbar, h->ubar, h->pf, h->A, h->Bm)
      if (h->seq_ok) { LIND_LAUNCH(0, NQ, true); LIND_LAUNCH(1, NV, true); LIND_LAUNCH(2, NU, true); }
      else { LIND_LAUNCH(0, NQ, false); LIND_LAUNCH(1, NV, false); LIND_LAUNCH(2, NU, false); }
#undef LIND_LAUNCH
      h->launches += 3;
      return;
    }
    k_linearize_analytic<<<h->B * h->N, LINA_WARPS * 32, h->smem_lina, h->stream>>>(h->d_dyn, h->N, mask, h->xbar, h->ubar,
                                                                                 h->pf, h->A, h->Bm);
    LAUNCHED();
    return;
  }
  k_linearize_fd<<<h->B * h->N, LIN_WARPS * 32, h->smem_lin, h->stream>>>(h->d_dyn, h->N, h->opt.fd_eps, mask, h->xbar,
                                                                         h->ubar, h->A, h->Bm);
  LAUNCHED();
}
static void launch_cost_quadratics(H1Ilqr* h, const int* mask) {
  const long warps = (long)h->B * (h->N + 1);   // knots (two warps each)
  const int blocks = (int)((warps + CQ_KNOTS - 1) / CQ_KNOTS);
  if (h->qoff_on)
    k_cost_quadratics<true><<<blocks, CQ_WARPS * 32, h->smem_cq, h->stream>>>(h->d_cost, h->d_dyn, h->d_w, ref_table(h), h->B, h->N, mask, h->xbar,
                                                                             h->ubar, h->lx, h->lu, h->lxx, h->luu, h->d_qoff);
  else
    k_cost_quadratics<false><<<blocks, CQ_WARPS * 32, h->smem_cq, h->stream>>>(h->d_cost, h->d_dyn, h->d_w, ref_table(h), h->B, h->N, mask, h->xbar,
                                                                              h->ubar, h->lx, h->lu, h->lxx, h->luu, nullptr);
  LAUNCHED();
}
static void launch_backward(H1Ilqr* h, const int* mask) {
  if (h->ab_struct && h->ric_struct_ok)
    k_backward<true><<<h->B, RIC_THREADS, h->smem_ric, h->stream>>>(h->N, mask, h->lambda, h->A, h->Bm, h->lx, h->lu, h->lxx,
                                                                   h->luu, h->K, h->kff, h->status, h->dt);
  else
    k_backward<false><<<h->B, RIC_THREADS, h->smem_ric, h->stream>>>(h->N, mask, h->lambda, h->A, h->Bm, h->lx, h->lu, h->lxx,
                                                                    h->luu, h->K, h->kff, h->status, 0.0);
  LAUNCHED();
}
static void launch_line_search(H1Ilqr* h, const int* mask) {
  if (h->seq_ok && h->ls_quad && h->policy != H1ILQR_KERNELS_COOPERATIVE) {   // one warp per instance: 8 candidates x 4 chains (any batch size)
    const int* cnt; const int* list = list_for(h, mask, &cnt);
#define Q4_LAUNCH(W)                                                                                            \
  k_line_search_quad<W><<<(unsigned)((h->B + W - 1) / W), W * 32, h->smem_q4, h->stream>>>(                     \
      h->d_dyn, h->d_w, h->d_opt, ref_table(h), h->B, h->N, mask, list, cnt, h->x0, h->nominal_cost, h->xbar, h->ubar, \
      h->K, h->kff, h->xnew, h->unew, h->ls_ok, h->ls_cost, h->ls_alpha)
    if (h->q4_warps == 1) Q4_LAUNCH(1); else Q4_LAUNCH(8);
#undef Q4_LAUNCH
    LAUNCHED();
    return;
  }
  if (h->seq_ok && use_batched(h, h->B, h->seq_min_batch)) {   // one thread per (instance, candidate)
    const long threads = (long)h->B * H1ILQR_NALPHA;
    const int* cnt; const int* list = list_for(h, mask, &cnt);
    k_line_search_seq<<<(unsigned)((threads + SEQ_THREADS - 1) / SEQ_THREADS), SEQ_THREADS, h->smem_seq_ls, h->stream>>>(
        h->d_dyn, h->d_w, h->d_opt, ref_table(h), h->B, h->N, mask, list, cnt,
        h->x0, h->nominal_cost, h->xbar, h->ubar, h->K, h->kff,
        h->xnew, h->unew, h->ls_ok, h->ls_cost, h->ls_alpha);
    LAUNCHED();
    return;
  }
  k_line_search<<<h->B, H1ILQR_NALPHA * 32, h->smem_ls, h->stream>>>(h->d_dyn, h->d_w, h->d_opt, ref_table(h), h->N, mask,
                                                                    h->x0, h->nominal_cost, h->xbar, h->ubar, h->K, h->kff,
                                                                    h->xnew, h->unew, h->ls_ok, h->ls_cost, h->ls_alpha);
  LAUNCHED();
}
static SolveState solve_state(H1Ilqr* h) {
  SolveState st;
  st.lambda = h->lambda; st.cost = h->cost; st.prev_cost = h->prev_cost; st.nominal_cost = h->nominal_cost;
  st.active = h->active; st.second = h->second; st.iters = h->iters; st.status = h->status;
  st.ls_ok = h->ls_ok; st.ls_cost = h->ls_cost; st.ls_alpha = h->ls_alpha;
  st.cost_trace = h->cost_trace; st.alpha_trace = h->alpha_trace;
  st.act_list = h->act_list; st.sec_list = h->sec_list; st.early_list = h->early_list; st.late_list = h->late_list;
  st.early = h->early; st.late = h->late; st.list_count = h->list_count;
  return st;
}
static void launch_state(H1Ilqr* h, int it, int phase) {
  // counters of the lists this phase builds: 0 -> active; 1 -> second, early; 2 -> late
  if (phase == 0) cudaMemsetAsync(h->list_count, 0, sizeof(int), h->stream);
  else if (phase == 1) cudaMemsetAsync(h->list_count + 1, 0, 2 * sizeof(int), h->stream);
  else cudaMemsetAsync(h->list_count + 3, 0, sizeof(int), h->stream);
  k_solve_state<<<(h->B + 127) / 128, 128, 0, h->stream>>>(solve_state(h), h->d_opt, h->B, it, phase);
  LAUNCHED();
}

struct StageTimer {
  H1Ilqr* h; double* acc;
  StageTimer(H1Ilqr* h_, double* acc_) : h(h_), acc(acc_) { if (h->timing) cudaEventRecord(h->ev[0], h->stream); }
  ~StageTimer() {
    if (!h->timing) return;
    cudaEventRecord(h->ev[1], h->stream);
    cudaEventSynchronize(h->ev[1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    *acc += ms;
  }
};

// The whole iLQR::solve launch sequence, stream-ordered, no host round trips (unless stage timing is on).
//
// iLQR::forwardRolloutNominal (ilqr.cpp:551-563): in iteration 0 the guess is rolled out from x0. In later iterations
// xbar/ubar are what the last accepted line-search candidate left (or unchanged after a failed one): rolling them out
// again from the same x0 reproduces them (f_D is deterministic), so the sequential rollout is replaced by the
// knot-parallel factorisation and the known cost (k_solve_state, phase 0).
//
// Pipelining: after the first line search of iteration `it` the instances that accepted a candidate are done with the
// iteration (k_solve_state phase 1), so factorisation + linearization + cost quadratics of iteration it + 1 start
// for them at once on the main stream, while backward pass + line search of the second attempts — a small, latency-
// bound set — run on `stream2`. Second attempts that succeed (rare) are linearized after the join; instances whose
// two attempts both failed keep their trajectory and therefore their A, B, lx, lxx ... of this iteration.
static void enqueue_solve(H1Ilqr* h) {
  const bool analytic = h->opt.linearization != H1ILQR_LIN_FD;
  cudaStream_t main_stream = h->stream;
  auto derivatives = [&](const int* mask) {   // of the current trajectory of the instances in `mask`
    if (analytic) { StageTimer t(h, &h->times.rollout_ms); launch_factors(h, mask); }
    { StageTimer t(h, &h->times.linearize_ms); launch_linearize(h, mask, analytic); }
    { StageTimer t(h, &h->times.cost_quadratics_ms); launch_cost_quadratics(h, mask); }
  };
  launch_rollout(h, nullptr, nullptr, h->N, h->cost);  // current_cost = computeTotalCost(xbar, ubar)
  for (int it = 0; it < h->opt.max_iterations; ++it) {
    launch_state(h, it, 0);
    if (it == 0) {
      { StageTimer t(h, &h->times.rollout_ms); launch_rollout(h, h->active, h->x0, 0, h->nominal_cost, analytic); }
      { StageTimer t(h, &h->times.linearize_ms); launch_linearize(h, h->active, analytic); }
      { StageTimer t(h, &h->times.cost_quadratics_ms); launch_cost_quadratics(h, h->active); }
    }
    { StageTimer t(h, &h->times.backward_ms); launch_backward(h, h->active); }
    { StageTimer t(h, &h->times.line_search_ms); launch_line_search(h, h->active); }
    launch_state(h, it, 1);
    const bool more = it + 1 < h->opt.max_iterations;
    // fork: second attempts on stream2 ...
    cudaEventRecord(h->ev_fork, main_stream);
    cudaStreamWaitEvent(h->stream2, h->ev_fork, 0);
    h->stream = h->stream2;
    { StageTimer t(h, &h->times.backward_ms); launch_backward(h, h->second); }
    { StageTimer t(h, &h->times.line_search_ms); launch_line_search(h, h->second); }
    launch_state(h, it, 2);
    cudaEventRecord(h->ev_join, h->stream2);
    h->stream = main_stream;
    // ... next iteration's derivatives of the early set on the main stream, then join and the late set
    if (more) derivatives(h->early);
    cudaStreamWaitEvent(main_stream, h->ev_join, 0);
    if (more) derivatives(h->late);
  }
}

static int enqueue_initialize(H1Ilqr* h, const int* warm_dev, int u_shared) {
  k_init_guess<<<h->B, 128, 0, h->stream>>>(h->B, h->N, h->x0, warm_dev, h->has_prev, h->u_init, u_shared, h->prev_xbar,
                                            h->prev_ubar, h->xbar, h->ubar, h->warm_mask, h->cold_mask);
  LAUNCHED();
  launch_rollout(h, h->cold_mask, nullptr, 0, nullptr);
  launch_rollout(h, h->warm_mask, nullptr, h->N - 1, nullptr);
  return 0;
}

int h1ilqr_initialize(H1Ilqr* h, const double* x0, const int* warm, const double* u_init, int u_init_shared) {
  GUARD(h);
  if (!x0) return set_err(H1ILQR_EARG, "null x0");
  H2D(h->x0, x0, (size_t)h->B * NX * sizeof(double));
  if (u_init) H2D(h->u_init, u_init, (u_init_shared ? 1 : (size_t)h->B) * NU * sizeof(double));
  else CU(cudaMemsetAsync(h->u_init, 0, (size_t)h->B * NU * sizeof(double), h->stream));
  const int* wd = nullptr;
  if (warm) { H2D(h->warm_in, warm, (size_t)h->B * sizeof(int)); wd = h->warm_in; }
  else { k_fill_int<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->warm_in, 0); wd = h->warm_in; }
  enqueue_initialize(h, wd, u_init ? u_init_shared : 1);
  SYNC();
  CU(cudaGetLastError());
  return 0;
}

static int finish_solve(H1Ilqr* h, double* cost_out, int* iters_out, int* status_out) {
  if (cost_out) D2H(cost_out, h->cost, (size_t)h->B * sizeof(double));
  if (iters_out) D2H(iters_out, h->iters, (size_t)h->B * sizeof(int));
  std::vector<int> st(h->B);
  D2H(st.data(), h->status, (size_t)h->B * sizeof(int));
  SYNC();
  CU(cudaGetLastError());
  int bad = 0;
  for (int i = 0; i < h->B; ++i) bad |= st[i];
  if (status_out) std::memcpy(status_out, st.data(), (size_t)h->B * sizeof(int));
  return bad ? H1ILQR_ENOTFINITE : 0;
}

int h1ilqr_solve(H1Ilqr* h, const double* x0, double* cost_out, int* iters_out, int* status_out) {
  GUARD(h);
  if (!x0) return set_err(H1ILQR_EARG, "null x0");
  H2D(h->x0, x0, (size_t)h->B * NX * sizeof(double));
  const int l0 = h->launches;
  std::memset(&h->times, 0, sizeof(h->times));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, h->stream));
  enqueue_solve(h);
  CU(cudaEventRecord(e1, h->stream));
  int rc = finish_solve(h, cost_out, iters_out, status_out);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  h->times.total_ms = ms; h->times.launches = h->launches - l0;
  if (rc == H1ILQR_ENOTFINITE) set_err(rc, "non-finite cost or gains in at least one instance");
  return rc;
}

static void enqueue_mpc_tail(H1Ilqr* h) {
  const size_t B = h->B, N = h->N;
  k_first_control<<<h->B, 32, 0, h->stream>>>(h->B, h->N, h->x0, h->xbar, h->ubar, h->K, h->u_apply);
  LAUNCHED();
  cudaMemcpyAsync(h->prev_xbar, h->xbar, B * (N + 1) * NX * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
  cudaMemcpyAsync(h->prev_ubar, h->ubar, B * N * NU * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
  k_fill_int<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->has_prev, 1);
  LAUNCHED();
}

int h1ilqr_host_register(H1Ilqr* h, const void* host_ptr, size_t bytes) {
  GUARD(h);
  if (!host_ptr || bytes == 0) return set_err(H1ILQR_EARG, "h1ilqr_host_register: null buffer");
  CU(cudaHostRegister(const_cast<void*>(host_ptr), bytes, cudaHostRegisterDefault));
  return 0;
}
int h1ilqr_host_unregister(H1Ilqr* h, const void* host_ptr) {
  GUARD(h);
  if (!host_ptr) return set_err(H1ILQR_EARG, "h1ilqr_host_unregister: null buffer");
  CU(cudaHostUnregister(const_cast<void*>(host_ptr)));
  return 0;
}

int h1ilqr_upload_inputs(H1Ilqr* h, const double* x_measured, const double* u_init, int u_init_shared) {
  GUARD(h);
  if (!x_measured) return set_err(H1ILQR_EARG, "null x_measured");
  H2D(h->x0, x_measured, (size_t)h->B * NX * sizeof(double));
  if (u_init) {
    if (u_init_shared) {  // replicate so that the resident steps can always read per-instance guesses
      std::vector<double> rep((size_t)h->B * NU);
      for (int i = 0; i < h->B; ++i) std::memcpy(&rep[(size_t)i * NU], u_init, NU * sizeof(double));
      H2D(h->u_init, rep.data(), rep.size() * sizeof(double));
      SYNC();
    } else H2D(h->u_init, u_init, (size_t)h->B * NU * sizeof(double));
  } else CU(cudaMemsetAsync(h->u_init, 0, (size_t)h->B * NU * sizeof(double), h->stream));
  SYNC();
  return 0;
}

// ---- one MPC step as a launch sequence (directly on the stream, or captured once into a CUDA graph) ----
static void enqueue_resident_step(H1Ilqr* h, bool cold) {
  if (cold) {
    k_fill_int<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->has_prev, 0);
    k_fill_double<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->lambda, h->opt.reg_init);
    h->launches += 2;
  }
  enqueue_initialize(h, nullptr, 0);
  enqueue_solve(h);
  enqueue_mpc_tail(h);
}
static RefTables ref_tables(const H1Ilqr* h) {
  RefTables tb;
  tb.x = h->tab_x; tb.com = h->tab_com; tb.ee = h->tab_ee; tb.cv = h->tab_cv; tb.contact = h->tab_contact;
  tb.rows = h->tab_rows; tb.contact_rows = h->tab_contact_rows; tb.schedule_offset = h->tab_schedule_offset;
  return tb;
}
// closed-loop step: window at t_idx -> MPC step (warm start from the previous solution) -> plant x <- f_D(x, u_apply) -> t_idx++
static void enqueue_closed_loop_step(H1Ilqr* h, bool logs) {
  const long nw = (long)h->B * (h->N + 1);
  k_extract_window<<<(unsigned)((nw + 127) / 128), 128, 0, h->stream>>>(ref_tables(h), h->B, h->N, h->t_idx, h->x_ref, h->com_ref,
                                                                      h->ee_ref, h->com_vel_ref, h->stance);
  LAUNCHED();
  enqueue_resident_step(h, false);
  k_dyn_step<<<(h->B + 3) / 4, 128, h->smem_dyn4, h->stream>>>(h->d_dyn, h->B, h->x0, h->u_apply, h->plant_next);
  cudaMemcpyAsync(h->x0, h->plant_next, (size_t)h->B * NX * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
  k_closed_loop_advance<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->t_idx, h->step_ctr, h->cost, h->iters, h->u_apply,
                                                                  logs ? h->log_cost : nullptr, logs ? h->log_iters : nullptr,
                                                                  logs ? h->log_u : nullptr);
  k_increment<<<1, 1, 0, h->stream>>>(h->step_ctr);
  h->launches += 3;
}
// Capture `which` (0 resident warm, 1 resident cold, 2 closed loop, 3 closed loop with logs) once; the launch sequence of a
// step is the same for every step (iteration counts are handled on the device by the compact instance lists), so the
// graph is replayed unchanged. Both streams of the pipelined solve are part of the capture (fork / join events).
static int ensure_graph(H1Ilqr* h, int which) {
  if (h->graph[which]) return 0;
  const int l0 = h->launches;
  cudaGraph_t g = nullptr;
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  if (which < 2) enqueue_resident_step(h, which == 1); else enqueue_closed_loop_step(h, which == 3);
  cudaError_t e = cudaStreamEndCapture(h->stream, &g);
  h->graph_launches[which] = h->launches - l0;
  h->launches = l0;
  if (e != cudaSuccess) return set_err(H1ILQR_ECUDA, "cudaStreamEndCapture", e);
  e = cudaGraphInstantiate(&h->graph[which], g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { h->graph[which] = nullptr; return set_err(H1ILQR_ECUDA, "cudaGraphInstantiate", e); }
  return 0;
}

int h1ilqr_set_reference_table(H1Ilqr* h, int rows, const double* x_ref_full, const double* com_ref_full,
                               const double* ee_ref_full, const double* com_vel_ref_full, int contact_rows,
                               const int* contact, int schedule_offset) {
  GUARD(h);
  if (rows < 1 || !x_ref_full || !com_ref_full || !ee_ref_full || contact_rows < 0 || (contact_rows > 0 && !contact))
    return set_err(H1ILQR_EARG, "h1ilqr_set_reference_table: bad arguments");
  // getEEReference / getCoMVelReference throw past the table (robot_utils.cpp:525-549): the horizon-local lookups need N + 1 rows
  if (!schedule_offset && rows < h->N + 1) return set_err(H1ILQR_EARG, "Invalid reference index: the table has fewer than N + 1 rows");
  invalidate_graphs(h);
  for (void* p : h->table_allocs) cudaFree(p);
  h->table_allocs.clear();
  auto up = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes ? bytes : 8);
    if (e != cudaSuccess) return e;
    h->table_allocs.push_back(*dst);
    if (src) return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, h->stream);
    return cudaMemsetAsync(*dst, 0, bytes ? bytes : 8, h->stream);
  };
  CU(up(x_ref_full, (size_t)rows * NX * sizeof(double), (void**)&h->tab_x));
  CU(up(com_ref_full, (size_t)rows * 3 * sizeof(double), (void**)&h->tab_com));
  CU(up(ee_ref_full, (size_t)rows * 6 * sizeof(double), (void**)&h->tab_ee));
  CU(up(com_vel_ref_full, (size_t)rows * 3 * sizeof(double), (void**)&h->tab_cv));   // NULL: zeros (only read when W_com_vel > 0)
  CU(up(contact, (size_t)contact_rows * 2 * sizeof(int), (void**)&h->tab_contact));
  h->tab_rows = rows; h->tab_contact_rows = contact_rows; h->tab_schedule_offset = schedule_offset ? 1 : 0;
  SYNC();
  return 0;
}

int h1ilqr_run_closed_loop(H1Ilqr* h, int steps, const int* t_idx0, const double* x_start, const double* u_init,
                           int u_init_shared, int use_graph, double* x_final, double* cost_log, int* iters_log,
                           double* u_log, double* elapsed_ms) {
  GUARD(h);
  if (steps < 1) return set_err(H1ILQR_EARG, "steps < 1");
  if (!h->tab_x) return set_err(H1ILQR_EARG, "h1ilqr_run_closed_loop: no reference table (h1ilqr_set_reference_table)");
  const size_t B = h->B;
  int rc = 0;
  if (x_start && (rc = h1ilqr_upload_inputs(h, x_start, u_init, u_init_shared))) return rc;
  if (t_idx0) H2D(h->t_idx, t_idx0, B * sizeof(int));
  const bool logs = cost_log || iters_log || u_log;
  if (logs && h->log_capacity < steps) {
    invalidate_graphs(h);
    if (h->log_cost) { cudaFree(h->log_cost); cudaFree(h->log_iters); cudaFree(h->log_u); }
    CU(cudaMalloc((void**)&h->log_cost, (size_t)steps * B * sizeof(double)));
    CU(cudaMalloc((void**)&h->log_iters, (size_t)steps * B * sizeof(int)));
    CU(cudaMalloc((void**)&h->log_u, (size_t)steps * B * NU * sizeof(double)));
    h->log_capacity = steps;
  }
  CU(cudaMemsetAsync(h->step_ctr, 0, sizeof(int), h->stream));
  CU(cudaMemsetAsync(h->u_ref, 0, B * h->N * NU * sizeof(double), h->stream));   // u_ref_full_ is all zeros (robot_utils.cpp:367)
  if (h->ref_shared) invalidate_graphs(h);
  h->ref_shared = 0;     // the loop writes per-instance windows
  const int which = logs ? 3 : 2;
  const bool timing = h->timing;
  h->timing = false;
  if (use_graph && (rc = ensure_graph(h, which))) { h->timing = timing; return rc; }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  const int l0 = h->launches;
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaEventRecord(e0, h->stream));
  for (int s = 0; s < steps; ++s) {
    if (use_graph) { CU(cudaGraphLaunch(h->graph[which], h->stream)); h->launches += h->graph_launches[which]; }
    else enqueue_closed_loop_step(h, logs);
  }
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  h->timing = timing;
  h->times.launches = h->launches - l0;
  if (elapsed_ms) *elapsed_ms = ms;
  if (x_final) D2H(x_final, h->x0, B * NX * sizeof(double));
  if (cost_log) D2H(cost_log, h->log_cost, (size_t)steps * B * sizeof(double));
  if (iters_log) D2H(iters_log, h->log_iters, (size_t)steps * B * sizeof(int));
  if (u_log) D2H(u_log, h->log_u, (size_t)steps * B * NU * sizeof(double));
  std::vector<int> st(B);
  D2H(st.data(), h->status, B * sizeof(int));
  SYNC();
  int bad = 0;
  for (size_t i = 0; i < B; ++i) bad |= st[i];
  if (bad) return set_err(H1ILQR_ENOTFINITE, "non-finite cost or gains in at least one instance at the last step");
  return 0;
}

int h1ilqr_run_resident_steps(H1Ilqr* h, int steps, int cold_each_step, double* elapsed_ms) {
  GUARD(h);
  if (steps < 1) return set_err(H1ILQR_EARG, "steps < 1");
  const bool use_graph = (cold_each_step & 2) != 0;   // bit 1: replay the step as a CUDA graph
  cold_each_step &= 1;
  const bool timing = h->timing;
  h->timing = false;  // stage timing would insert host syncs
  if (use_graph) { int rc = ensure_graph(h, cold_each_step); if (rc) { h->timing = timing; return rc; } }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  const int l0 = h->launches;
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaEventRecord(e0, h->stream));
  for (int s = 0; s < steps; ++s) {
    if (use_graph) { CU(cudaGraphLaunch(h->graph[cold_each_step], h->stream)); h->launches += h->graph_launches[cold_each_step]; }
    else enqueue_resident_step(h, cold_each_step != 0);
  }
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  h->timing = timing;
  h->times.launches = h->launches - l0;
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}
int h1ilqr_measure_fp64_peak(H1Ilqr* h, double* tflops) {
  GUARD(h);
  if (!tflops) return set_err(H1ILQR_EARG, "null tflops");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, h->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(e0, h->stream));
    k_fp64_peak<<<blocks, threads, 0, h->stream>>>(h->scratch, iters);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * threads * blocks;
    if (rep > 0) best = fmax(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return 0;
}

// fp64 tensor-core peak: mma.sync m8n8k4 (SASS DMMA), 8 independent accumulator chains per warp
__global__ void k_dmma_peak(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = 0.5 * i; }
  const double a = 1.0000001, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
int h1ilqr_measure_fp64_mma_peak(H1Ilqr* h, double* tflops) {
  GUARD(h);
  if (!tflops) return set_err(H1ILQR_EARG, "null tflops");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, h->device));
  const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 1 << 14;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(e0, h->stream));
    k_dmma_peak<<<blocks, threads, 0, h->stream>>>(h->scratch, iters);
    CU(cudaEventRecord(e1, h->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 256.0 * 8.0 * (double)iters * (threads / 32) * blocks;   // 8x8x4 FMAs per mma
    if (rep > 0) best = fmax(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return 0;
}

int h1ilqr_mpc_reset(H1Ilqr* h) {
  GUARD(h);
  k_fill_int<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->has_prev, 0);
  k_fill_double<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->lambda, h->opt.reg_init);
  SYNC();
  return 0;
}

int h1ilqr_mpc_step(H1Ilqr* h, const double* x_measured, const double* u_init, int u_init_shared, double* u_apply,
                    double* cost_out) {
  GUARD(h);
  if (!x_measured || !u_apply) return set_err(H1ILQR_EARG, "null x_measured / u_apply");
  const size_t B = h->B;
  // pinned staging: x in, u_apply + cost out
  double* pin_x = (double*)h->pin; double* pin_u = pin_x + B * NX; double* pin_c = pin_u + B * NU;
  double* pin_ui = pin_c + B;
  int* pin_st = reinterpret_cast<int*>(pin_ui + B * NU);
  std::memcpy(pin_x, x_measured, B * NX * sizeof(double));
  H2D(h->x0, pin_x, B * NX * sizeof(double));
  if (u_init) {
    const size_t n = (u_init_shared ? 1 : B) * NU;
    std::memcpy(pin_ui, u_init, n * sizeof(double));
    H2D(h->u_init, pin_ui, n * sizeof(double));
  } else CU(cudaMemsetAsync(h->u_init, 0, B * NU * sizeof(double), h->stream));
  const int l0 = h->launches;
  std::memset(&h->times, 0, sizeof(h->times));
  enqueue_initialize(h, nullptr, u_init ? u_init_shared : 1);
  enqueue_solve(h);
  enqueue_mpc_tail(h);
  D2H(pin_u, h->u_apply, B * NU * sizeof(double));
  D2H(pin_c, h->cost, B * sizeof(double));
  D2H(pin_st, h->status, B * sizeof(int));
  SYNC();
  CU(cudaGetLastError());
  h->times.launches = h->launches - l0;
  std::memcpy(u_apply, pin_u, B * NU * sizeof(double));
  if (cost_out) std::memcpy(cost_out, pin_c, B * sizeof(double));
  // same convention as h1ilqr_solve: the outputs are delivered, the return code says whether an instance went non-finite
  // (the reference carries non-finite gains on with a warning, ilqr.cpp:290-293; a batched caller must not miss it)
  int bad = 0;
  for (size_t i = 0; i < B; ++i) bad |= pin_st[i];
  if (bad) return set_err(H1ILQR_ENOTFINITE, "non-finite cost or gains in at least one instance (h1ilqr_get_status)");
  return 0;
}

int h1ilqr_get_status(H1Ilqr* h, int* status_out, int* iters_out) {
  GUARD(h);
  if (status_out) D2H(status_out, h->status, (size_t)h->B * sizeof(int));
  if (iters_out) D2H(iters_out, h->iters, (size_t)h->B * sizeof(int));
  SYNC();
  return 0;
}

// ---------------- granular stages ----------------
int h1ilqr_rollout_nominal(H1Ilqr* h, const double* x0) {
  GUARD(h);
  if (!x0) return set_err(H1ILQR_EARG, "null x0");
  H2D(h->x0, x0, (size_t)h->B * NX * sizeof(double));
  launch_rollout(h, nullptr, h->x0, 0, h->nominal_cost);
  SYNC(); CU(cudaGetLastError());
  return 0;
}
int h1ilqr_linearize(H1Ilqr* h) { GUARD(h); launch_linearize(h, nullptr); SYNC(); CU(cudaGetLastError()); return 0; }
int h1ilqr_cost_quadratics(H1Ilqr* h) { GUARD(h); launch_cost_quadratics(h, nullptr); SYNC(); CU(cudaGetLastError()); return 0; }
#ifdef LINF_PROF   // debug build only: read and clear the per-step cycle sums of k_linearize_finish (not part of the ABI)
int h1ilqr_debug_linf_prof(unsigned long long* out8) {
  CU(cudaDeviceSynchronize());
  static const unsigned long long zeros[8] = {0};
  CU(cudaMemcpyFromSymbol(out8, h1::linf_prof_sum, sizeof(zeros)));
  CU(cudaMemcpyToSymbol(h1::linf_prof_sum, zeros, sizeof(zeros)));
  return 0;
}
#endif
#ifdef CQ_PROF   // debug build only: read and clear the per-phase cycle sums of k_cost_quadratics (not part of the ABI)
int h1ilqr_debug_cq_prof(unsigned long long* out64) {
  CU(cudaDeviceSynchronize());
  static const unsigned long long zeros[64] = {0};
  CU(cudaMemcpyFromSymbol(out64, h1::cq_prof_sum, sizeof(zeros)));
  CU(cudaMemcpyToSymbol(h1::cq_prof_sum, zeros, sizeof(zeros)));
  return 0;
}
#endif
#ifdef LINT_PROF   // debug build only: read and clear the per-item cycle sums of k_linearize_tangents (not part of the ABI)
int h1ilqr_debug_lint_prof(unsigned long long* out131) {
  CU(cudaDeviceSynchronize());
  static const unsigned long long zeros[128] = {0};
  CU(cudaMemcpyFromSymbol(out131, h1::lint_prof_item, sizeof(unsigned long long) * 128));
  CU(cudaMemcpyFromSymbol(out131 + 128, h1::lint_prof_busy, sizeof(unsigned long long)));
  CU(cudaMemcpyFromSymbol(out131 + 129, h1::lint_prof_span, sizeof(unsigned long long)));
  CU(cudaMemcpyToSymbol(h1::lint_prof_item, zeros, sizeof(zeros)));
  CU(cudaMemcpyToSymbol(h1::lint_prof_busy, zeros, sizeof(unsigned long long)));
  CU(cudaMemcpyToSymbol(h1::lint_prof_span, zeros, sizeof(unsigned long long)));
  return 0;
}
#endif
#ifdef RIC_PROF   // debug build only: read and clear the per-phase cycle sums of k_backward (not part of the ABI)
int h1ilqr_debug_ric_prof(unsigned long long* out40) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out40, h1::ric_prof_sum, sizeof(unsigned long long) * 40));
  static const unsigned long long zeros[40] = {0};
  CU(cudaMemcpyToSymbol(h1::ric_prof_sum, zeros, sizeof(zeros)));
  return 0;
}
#endif
int h1ilqr_backward_pass(H1Ilqr* h) { GUARD(h); launch_backward(h, nullptr); SYNC(); CU(cudaGetLastError()); return 0; }
int h1ilqr_total_cost(H1Ilqr* h, double* cost_out) {
  GUARD(h);
  if (!cost_out) return set_err(H1ILQR_EARG, "null cost_out");
  launch_rollout(h, nullptr, nullptr, h->N, h->nominal_cost);
  D2H(cost_out, h->nominal_cost, (size_t)h->B * sizeof(double));
  SYNC(); CU(cudaGetLastError());
  return 0;
}
int h1ilqr_line_search(H1Ilqr* h, const double* x0, int* improved, double* new_cost, int* alpha_index) {
  GUARD(h);
  if (!x0) return set_err(H1ILQR_EARG, "null x0");
  H2D(h->x0, x0, (size_t)h->B * NX * sizeof(double));  // candidates start from x0 (ilqr.cpp:327)
  launch_rollout(h, nullptr, nullptr, h->N, h->nominal_cost);  // baseline = computeTotalCost(xbar, ubar)
  launch_line_search(h, nullptr);
  if (improved) D2H(improved, h->ls_ok, (size_t)h->B * sizeof(int));
  if (new_cost) D2H(new_cost, h->ls_cost, (size_t)h->B * sizeof(double));
  if (alpha_index) D2H(alpha_index, h->ls_alpha, (size_t)h->B * sizeof(int));
  SYNC(); CU(cudaGetLastError());
  return 0;
}

// Device time of `reps` back-to-back launches of one stage on the current trajectory / derivatives / gains of every
// instance (CUDA events on the handle's stream): the per-kernel launch duration the roofline figures are built from.
// stage: 0 factorisation of Mhat (analytic linearization input), 1 linearization (factors ready), 2 cost quadratics,
// 3 backward pass, 4 line search (candidates are rolled out, the accepted one is NOT installed: xbar / ubar are restored).
int h1ilqr_time_stage(H1Ilqr* h, int stage, int reps, double* elapsed_ms) {
  GUARD(h);
  if (stage < 0 || stage > 4 || reps < 1 || !elapsed_ms) return set_err(H1ILQR_EARG, "h1ilqr_time_stage: bad arguments");
  const size_t B = h->B, N = h->N;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double *xs = nullptr, *us = nullptr;
  if (stage == 4) {   // the line search overwrites the trajectory of the instances that accept a candidate
    CU(cudaMalloc((void**)&xs, B * (N + 1) * NX * sizeof(double))); CU(cudaMalloc((void**)&us, B * N * NU * sizeof(double)));
    CU(cudaMemcpyAsync(xs, h->xbar, B * (N + 1) * NX * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(us, h->ubar, B * N * NU * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    k_copy_x0<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->N, h->xbar, h->x0);
    launch_rollout(h, nullptr, nullptr, h->N, h->nominal_cost);
  }
  float total = 0.f;
  for (int r = 0; r < reps; ++r) {
    CU(cudaEventRecord(e0, h->stream));
    switch (stage) {
      case 0: launch_factors(h, nullptr); break;
      case 1: launch_linearize(h, nullptr, h->opt.linearization != H1ILQR_LIN_FD); break;
      case 2: launch_cost_quadratics(h, nullptr); break;
      case 3: launch_backward(h, nullptr); break;
      default: launch_line_search(h, nullptr); break;
    }
    CU(cudaEventRecord(e1, h->stream));
    if (stage == 4) {
      CU(cudaMemcpyAsync(h->xbar, xs, B * (N + 1) * NX * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      CU(cudaMemcpyAsync(h->ubar, us, B * N * NU * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    total += ms;
  }
  SYNC(); CU(cudaGetLastError());
  if (xs) cudaFree(xs);
  if (us) cudaFree(us);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *elapsed_ms = total;
  return 0;
}

static int ensure_scratch(H1Ilqr* h, size_t bytes) {
  if (bytes <= h->scratch_bytes) return 0;
  double* p = nullptr;
  CU(cudaMalloc((void**)&p, bytes));
  h->allocs.push_back(p);
  h->scratch = p; h->scratch_bytes = bytes;
  return 0;
}

int h1ilqr_dynamics_step(H1Ilqr* h, int n, const double* x, const double* u, double* x_next) {
  GUARD(h);
  if (n < 1 || !x || !u || !x_next) return set_err(H1ILQR_EARG, "h1ilqr_dynamics_step: bad arguments");
  int rc = ensure_scratch(h, (size_t)n * (2 * NX + NU) * sizeof(double));
  if (rc) return rc;
  double* dx = h->scratch; double* du = dx + (size_t)n * NX; double* dn = du + (size_t)n * NU;
  H2D(dx, x, (size_t)n * NX * sizeof(double)); H2D(du, u, (size_t)n * NU * sizeof(double));
  k_dyn_step<<<(n + 3) / 4, 128, h->smem_dyn4, h->stream>>>(h->d_dyn, n, dx, du, dn);
  LAUNCHED();
  D2H(x_next, dn, (size_t)n * NX * sizeof(double));
  SYNC(); CU(cudaGetLastError());
  return 0;
}

static int query(H1Ilqr* h, int n, const double* x, double* bias, double* com, double* ee, double* sole = nullptr,
                 double* comvel = nullptr, double* eevel = nullptr) {
  if (n < 1 || !x) return set_err(H1ILQR_EARG, "bad query arguments");
  int rc = ensure_scratch(h, (size_t)n * (NX + NV + 18 + 3 * NCPT) * sizeof(double));
  if (rc) return rc;
  double* dx = h->scratch; double* db = dx + (size_t)n * NX; double* dc = db + (size_t)n * NV; double* de = dc + (size_t)n * 3;
  double* ds = de + (size_t)n * 6; double* dv = ds + (size_t)n * 3 * NCPT; double* dw = dv + (size_t)n * 3;
  H2D(dx, x, (size_t)n * NX * sizeof(double));
  k_dyn_query<<<(n + 3) / 4, 128, h->smem_dyn4, h->stream>>>(h->d_dyn, n, dx, bias ? db : nullptr, com ? dc : nullptr,
                                                            ee ? de : nullptr, sole ? ds : nullptr, comvel ? dv : nullptr,
                                                            eevel ? dw : nullptr);
  LAUNCHED();
  if (comvel) D2H(comvel, dv, (size_t)n * 3 * sizeof(double));
  if (eevel) D2H(eevel, dw, (size_t)n * 6 * sizeof(double));
  if (bias) D2H(bias, db, (size_t)n * NV * sizeof(double));
  if (com) D2H(com, dc, (size_t)n * 3 * sizeof(double));
  if (ee) D2H(ee, de, (size_t)n * 6 * sizeof(double));
  if (sole) D2H(sole, ds, (size_t)n * 3 * NCPT * sizeof(double));
  SYNC(); CU(cudaGetLastError());
  return 0;
}
int h1ilqr_bias_forces(H1Ilqr* h, int n, const double* x, double* bias) { GUARD(h); return query(h, n, x, bias, nullptr, nullptr); }
int h1ilqr_reference_kinematics(H1Ilqr* h, int n, const double* x, double* com, double* ee) { GUARD(h); return query(h, n, x, nullptr, com, ee); }
int h1ilqr_reference_com_velocity(H1Ilqr* h, int n, const double* x, double* com_vel) {
  GUARD(h);
  if (!com_vel) return set_err(H1ILQR_EARG, "null com_vel");
  return query(h, n, x, nullptr, nullptr, nullptr, nullptr, com_vel);
}
int h1ilqr_reference_ee_velocity(H1Ilqr* h, int n, const double* x, double* ee_vel) {
  GUARD(h);
  if (!ee_vel) return set_err(H1ILQR_EARG, "null ee_vel");
  return query(h, n, x, nullptr, nullptr, nullptr, nullptr, nullptr, ee_vel);
}

// One (x, u) pair linearized on its own: RobotUtils::linearizeDynamicsFD (robot_utils.cpp:120-160) when mode ==
// H1ILQR_LIN_FD (forward differences with `eps`), the exact Jacobians of f_D otherwise. Works on scratch buffers: the
// solver state of the handle is untouched. A [51 x 51], B [51 x 19] column-major.
int h1ilqr_linearize_state(H1Ilqr* h, int mode, double eps, const double* x, const double* u, double* A, double* B) {
  GUARD(h);
  if (!x || !u || !A || !B) return set_err(H1ILQR_EARG, "h1ilqr_linearize_state: null argument");
  const size_t need = (size_t)(2 * NX + NU + NX * NX + NX * NU) * sizeof(double) + sizeof(PrimalFactor);
  int rc = ensure_scratch(h, need);
  if (rc) return rc;
  double* dx = h->scratch; double* du = dx + 2 * NX; double* dA = du + NU; double* dB = dA + NX * NX;
  PrimalFactor* pf = reinterpret_cast<PrimalFactor*>(dB + NX * NU);
  H2D(dx, x, NX * sizeof(double)); H2D(du, u, NU * sizeof(double));
  if (mode == H1ILQR_LIN_FD) {
    if (!(eps > 0.0)) return set_err(H1ILQR_EARG, "h1ilqr_linearize_state: eps must be positive");
    k_linearize_fd<<<1, LIN_WARPS * 32, h->smem_lin, h->stream>>>(h->d_dyn, 1, eps, nullptr, dx, du, dA, dB);
  } else {
    k_primal_factor<<<1, 128, h->smem_dyn4, h->stream>>>(h->d_dyn, 1, 1, nullptr, dx, du, pf);
    k_linearize_analytic<<<1, LINA_WARPS * 32, h->smem_lina, h->stream>>>(h->d_dyn, 1, nullptr, dx, du, pf, dA, dB);
  }
  h->launches += 2;
  D2H(A, dA, (size_t)NX * NX * sizeof(double)); D2H(B, dB, (size_t)NX * NU * sizeof(double));
  SYNC(); CU(cudaGetLastError());
  return 0;
}

// RobotUtils::constraintCost / constraintGradients / constraintHessians (robot_utils.cpp:615-778) for n (x, u) pairs.
// u == NULL: joint-limit terms only. Outputs (any may be NULL): cost [n], grad_x [n][51], grad_u [n][19], and the DIAGONALS
// of the (diagonal) Hessians hess_xx [n][51], hess_uu [n][19].
int h1ilqr_limit_penalties(H1Ilqr* h, int n, const double* x, const double* u, double* cost, double* grad_x, double* grad_u,
                           double* hess_xx_diag, double* hess_uu_diag) {
  GUARD(h);
  if (n < 1 || !x) return set_err(H1ILQR_EARG, "h1ilqr_limit_penalties: bad arguments");
  int rc = ensure_scratch(h, (size_t)n * (3 * NX + 3 * NU + 1) * sizeof(double));
  if (rc) return rc;
  double* dx = h->scratch; double* du = dx + (size_t)n * NX; double* dc = du + (size_t)n * NU; double* dgx = dc + n;
  double* dgu = dgx + (size_t)n * NX; double* dhx = dgu + (size_t)n * NU; double* dhu = dhx + (size_t)n * NX;
  H2D(dx, x, (size_t)n * NX * sizeof(double));
  if (u) H2D(du, u, (size_t)n * NU * sizeof(double));
  CU(cudaMemsetAsync(dgu, 0, (size_t)n * NU * sizeof(double), h->stream));
  CU(cudaMemsetAsync(dhu, 0, (size_t)n * NU * sizeof(double), h->stream));
  k_limit_penalties<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_dyn, h->d_w, n, dx, u ? du : nullptr, dc, dgx, dgu, dhx, dhu);
  LAUNCHED();
  if (cost) D2H(cost, dc, (size_t)n * sizeof(double));
  if (grad_x) D2H(grad_x, dgx, (size_t)n * NX * sizeof(double));
  if (grad_u) D2H(grad_u, dgu, (size_t)n * NU * sizeof(double));
  if (hess_xx_diag) D2H(hess_xx_diag, dhx, (size_t)n * NX * sizeof(double));
  if (hess_uu_diag) D2H(hess_uu_diag, dhu, (size_t)n * NU * sizeof(double));
  SYNC(); CU(cudaGetLastError());
  return 0;
}

// RobotUtils::stageCost (u != NULL) / terminalCost (u == NULL) (robot_utils.cpp:162-252) of n states against the given
// reference rows x_ref [n][51], u_ref [n][19] (NULL = zeros), com_ref [n][3] (NULL: no CoM term); cost [n].
int h1ilqr_stage_cost(H1Ilqr* h, int n, const double* x, const double* u, const double* x_ref, const double* u_ref,
                      const double* com_ref, double* cost) {
  GUARD(h);
  if (n < 1 || !x || !x_ref || !cost) return set_err(H1ILQR_EARG, "h1ilqr_stage_cost: bad arguments");
  int rc = ensure_scratch(h, (size_t)n * (2 * NX + 2 * NU + 4) * sizeof(double));
  if (rc) return rc;
  double* dx = h->scratch; double* du = dx + (size_t)n * NX; double* dxr = du + (size_t)n * NU; double* dur = dxr + (size_t)n * NX;
  double* dcr = dur + (size_t)n * NU; double* dc = dcr + (size_t)n * 3;
  H2D(dx, x, (size_t)n * NX * sizeof(double)); H2D(dxr, x_ref, (size_t)n * NX * sizeof(double));
  if (u) H2D(du, u, (size_t)n * NU * sizeof(double));
  if (u_ref) H2D(dur, u_ref, (size_t)n * NU * sizeof(double));
  if (com_ref) H2D(dcr, com_ref, (size_t)n * 3 * sizeof(double));
  k_stage_cost<<<(n + 3) / 4, 128, h->smem_dyn4, h->stream>>>(h->d_dyn, h->d_w, n, dx, u ? du : nullptr, dxr, u_ref ? dur : nullptr,
                                                             com_ref ? dcr : nullptr, dc);
  LAUNCHED();
  D2H(cost, dc, (size_t)n * sizeof(double));
  SYNC(); CU(cudaGetLastError());
  return 0;
}

int h1ilqr_sole_points(H1Ilqr* h, int n, const double* x, double* pts) {
  GUARD(h);
  if (!pts) return set_err(H1ILQR_EARG, "null pts");
  return query(h, n, x, nullptr, nullptr, nullptr, pts);
}

// ---------------- accessors ----------------
#define COPY_PAIR(fn_get, fn_set, p1, n1, p2, n2)                                              \
  int fn_get(H1Ilqr* h, double* a, double* b) {                                                \
    GUARD(h);                                                                                  \
    if (a) D2H(a, h->p1, (n1) * sizeof(double));                                               \
    if (b) D2H(b, h->p2, (n2) * sizeof(double));                                               \
    SYNC(); return 0;                                                                          \
  }                                                                                            \
  int fn_set(H1Ilqr* h, const double* a, const double* b) {                                    \
    GUARD(h);                                                                                  \
    if (a) H2D(h->p1, a, (n1) * sizeof(double));                                               \
    if (b) H2D(h->p2, b, (n2) * sizeof(double));                                               \
    SYNC(); return 0;                                                                          \
  }
#define SZ(x) ((size_t)h->B * (x))
COPY_PAIR(h1ilqr_get_trajectory, h1ilqr_set_trajectory, xbar, SZ((h->N + 1) * NX), ubar, SZ(h->N * NU))
COPY_PAIR(h1ilqr_get_gains, h1ilqr_set_gains, K, SZ(h->N * NU * NX), kff, SZ(h->N * NU))
// A_k, B_k, lxx_k sit at padded per-knot strides on the device (h1_common.cuh); the caller's arrays are dense
#define D2H_PITCH(dst, src, n, stride, count) CU(cudaMemcpy2DAsync(dst, (size_t)(n) * sizeof(double), src, (size_t)(stride) * sizeof(double), (size_t)(n) * sizeof(double), (size_t)(count), cudaMemcpyDeviceToHost, h->stream))
#define H2D_PITCH(dst, src, n, stride, count) CU(cudaMemcpy2DAsync(dst, (size_t)(stride) * sizeof(double), src, (size_t)(n) * sizeof(double), (size_t)(n) * sizeof(double), (size_t)(count), cudaMemcpyHostToDevice, h->stream))
int h1ilqr_get_linearization(H1Ilqr* h, double* A, double* B) {
  GUARD(h);
  if (A) D2H_PITCH(A, h->A, NX * NX, A_STRIDE, SZ(h->N));
  if (B) D2H_PITCH(B, h->Bm, NX * NU, B_STRIDE, SZ(h->N));
  SYNC(); return 0;
}
int h1ilqr_set_linearization(H1Ilqr* h, const double* A, const double* B) {
  GUARD(h);
  h->ab_struct = false;   // caller-supplied [A|B]: no structure assumed, dense Riccati contractions
  if (A) H2D_PITCH(h->A, A, NX * NX, A_STRIDE, SZ(h->N));
  if (B) H2D_PITCH(h->Bm, B, NX * NU, B_STRIDE, SZ(h->N));
  SYNC(); return 0;
}

// previous solution of the MPC loop (MPC::prev_xbar_ / prev_ubar_, mpc.hpp:57-58): the warm start shifts it (k_init_guess)
int h1ilqr_set_previous_solution(H1Ilqr* h, const double* prev_xbar, const double* prev_ubar) {
  GUARD(h);
  if (!prev_xbar || !prev_ubar) return set_err(H1ILQR_EARG, "null previous solution");
  H2D(h->prev_xbar, prev_xbar, SZ((h->N + 1) * NX) * sizeof(double));
  H2D(h->prev_ubar, prev_ubar, SZ(h->N * NU) * sizeof(double));
  k_fill_int<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->has_prev, 1);
  SYNC(); return 0;
}
int h1ilqr_get_previous_solution(H1Ilqr* h, double* prev_xbar, double* prev_ubar) {
  GUARD(h);
  if (prev_xbar) D2H(prev_xbar, h->prev_xbar, SZ((h->N + 1) * NX) * sizeof(double));
  if (prev_ubar) D2H(prev_ubar, h->prev_ubar, SZ(h->N * NU) * sizeof(double));
  SYNC(); return 0;
}

int h1ilqr_get_cost_quadratics(H1Ilqr* h, double* lx, double* lu, double* lxx, double* luu) {
  GUARD(h);
  if (lx) D2H(lx, h->lx, SZ((h->N + 1) * NX) * sizeof(double));
  if (lu) D2H(lu, h->lu, SZ(h->N * NU) * sizeof(double));
  if (lxx) {   // the device keeps the lower triangle: mirror it for the caller
    k_mirror_lower<<<(unsigned)SZ(h->N + 1), 128, 0, h->stream>>>((long)SZ(h->N + 1), h->lxx);
    LAUNCHED();
    D2H_PITCH(lxx, h->lxx, NX * NX, LXX_STRIDE, SZ(h->N + 1));
  }
  if (luu) D2H(luu, h->luu, SZ(h->N * NU * NU) * sizeof(double));
  SYNC(); return 0;
}
int h1ilqr_set_cost_quadratics(H1Ilqr* h, const double* lx, const double* lu, const double* lxx, const double* luu) {
  GUARD(h);
  if (lx) H2D(h->lx, lx, SZ((h->N + 1) * NX) * sizeof(double));
  if (lu) H2D(h->lu, lu, SZ(h->N * NU) * sizeof(double));
  if (lxx) H2D_PITCH(h->lxx, lxx, NX * NX, LXX_STRIDE, SZ(h->N + 1));
  if (luu) H2D(h->luu, luu, SZ(h->N * NU * NU) * sizeof(double));
  SYNC(); return 0;
}
int h1ilqr_get_regularization(H1Ilqr* h, double* lambda) {
  GUARD(h);
  if (!lambda) return set_err(H1ILQR_EARG, "null lambda");
  D2H(lambda, h->lambda, SZ(1) * sizeof(double)); SYNC(); return 0;
}
int h1ilqr_set_regularization(H1Ilqr* h, const double* lambda, int shared) {
  GUARD(h);
  if (!lambda) return set_err(H1ILQR_EARG, "null lambda");
  if (shared) { k_fill_double<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->lambda, lambda[0]); }
  else H2D(h->lambda, lambda, SZ(1) * sizeof(double));
  SYNC(); return 0;
}
int h1ilqr_get_solve_trace(H1Ilqr* h, double* cost_trace, int* alpha_trace) {
  GUARD(h);
  if (cost_trace) D2H(cost_trace, h->cost_trace, SZ(h->opt.max_iterations) * sizeof(double));
  if (alpha_trace) D2H(alpha_trace, h->alpha_trace, SZ(h->opt.max_iterations * 2) * sizeof(int));
  SYNC(); return 0;
}
int h1ilqr_set_kernel_policy(H1Ilqr* h, int policy) {
  GUARD(h);
  if (policy < H1ILQR_KERNELS_AUTO || policy > H1ILQR_KERNELS_BATCHED) return set_err(H1ILQR_EARG, "bad kernel policy");
  if (h->policy != policy) invalidate_graphs(h);
  h->policy = policy;
  return 0;
}
int h1ilqr_enable_stage_timing(H1Ilqr* h, int enable) { GUARD(h); h->timing = enable != 0; return 0; }
int h1ilqr_get_stage_times(H1Ilqr* h, H1StageTimes* t) {
  GUARD(h);
  if (!t) return set_err(H1ILQR_EARG, "null times");
  *t = h->times;
  return 0;
}

}  // extern "C"
