/*
 * h1ilqr.h — C ABI of the B200-native H1 iLQR solver core (libh1ilqr.so).
 *
 * This is the drop-in boundary for the hot path of premsuggu/mpc-ilqr-mujoco. The reference has no
 * FFI layer; its boundary is the C++ class API (include/ilqr/ilqr.hpp, include/ilqr/mpc.hpp,
 * include/common/robot_utils.hpp). The C++ shim classes in mpc-ilqr-mujoco_b200/host/ keep those
 * class names and signatures and call ONLY the functions below. Each entry point cites the
 * reference function it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - plain C, IEEE fp64 everywhere, caller-allocated flat HOST arrays unless a name ends in _dev;
 *  - matrices are COLUMN-MAJOR (Eigen's default) : A[k] is 51x51, B[k] 51x19, K[k] 19x51,
 *    lxx[k] 51x51, luu[k] 19x19;
 *  - batched arrays are instance-major: xbar is [batch][N+1][51], A is [batch][N][51*51] ...;
 *  - state layout x = [qpos(26); qvel(25)] in MuJoCo conventions (quaternion w,x,y,z; world-frame
 *    base linear velocity; body-frame base angular velocity), u = 19 motor torques;
 *  - every function returns 0 on success or a negative H1ILQR_E* code; no exceptions, no stdout;
 *  - one handle = one CUDA device context + one stream; a handle must not be used from two host
 *    threads at once (the reference is single-threaded, robot_utils.hpp:18).
 *  - there is NO CPU fallback: h1ilqr_create fails with H1ILQR_ECUDA when no sm_100 device exists.
 */
#ifndef H1ILQR_H
#define H1ILQR_H

#include "h1_model.h"

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define H1ILQR_OK 0
#define H1ILQR_EARG (-1)      /* bad argument (null pointer, size mismatch, bad stage) */
#define H1ILQR_ECUDA (-2)     /* CUDA runtime error; h1ilqr_last_error() has the text */
#define H1ILQR_ENOTFINITE (-3)/* a solve produced non-finite cost/gains for at least one instance */

#define H1ILQR_MAX_ITERS 64
#define H1ILQR_NALPHA 8
/* H1SolverOptions.linearization */
#define H1ILQR_LIN_ANALYTIC 0 /* exact dA/dB of f_D by forward-mode tangents (default; north star) */
#define H1ILQR_LIN_FD 1       /* the reference's method: forward differences, fd_eps (robot_utils.cpp:120-160) */

/* Cost weights. Reference: Config::buildCostMatrices (src/common/config.cpp:66-122) builds DIAGONAL
 * Q, R, Qf; the scalar task weights are RobotUtils::set*Weight (include/common/robot_utils.hpp:66-80)
 * and setConstraintWeights (src/common/robot_utils.cpp:674-680). */
typedef struct H1Weights {
  double Qdiag[H1_NX];
  double Rdiag[H1_NU];
  double Qfdiag[H1_NX];
  double w_com, w_com_vel, w_ee_pos, w_ee_vel, w_upright, w_balance;
  double w_joint_limits, w_control_limits;
} H1Weights;

/* Solver options. Reference defaults: iLQR::iLQR (src/ilqr/ilqr.cpp:14-16), lambda bounds
 * (ilqr.cpp:620,646), accept margin (ilqr.cpp:350), FD eps (include/common/robot_utils.hpp:53),
 * alpha list (ilqr.cpp:320). */
typedef struct H1SolverOptions {
  int max_iterations;      /* 10 */
  double tolerance;        /* 1e-4 */
  double reg_init;         /* 1e-6 */
  double reg_min;          /* 1e-6 */
  double reg_max;          /* 1e-3 */
  double accept_margin;    /* 1e-6 */
  double fd_eps;           /* 1e-5 */
  double divergence_cost;  /* 1e6 */
  int linearization;       /* H1ILQR_LIN_ANALYTIC (default) or H1ILQR_LIN_FD */
  double alphas[H1ILQR_NALPHA]; /* 1,.8,.6,.4,.2,.1,.05,.01 */
} H1SolverOptions;

void h1ilqr_default_options(H1SolverOptions* opt);

typedef struct H1Ilqr H1Ilqr; /* opaque */

/* Lifetime. `dyn_model` / `cost_model` may be NULL to use the built-in H1 tables.
 * Replaces: iLQR::iLQR buffer allocation (src/ilqr/ilqr.cpp:14-48) and MPC::MPC (src/ilqr/mpc.cpp:16-38),
 * for `batch` independent MPC instances resident on CUDA device `device`. */
int h1ilqr_create(const H1Model* dyn_model, const H1Model* cost_model, const H1SolverOptions* opt,
                  int batch, int N, int device, H1Ilqr** out);
void h1ilqr_destroy(H1Ilqr* h);
const char* h1ilqr_last_error(void);
int h1ilqr_batch(const H1Ilqr* h);
int h1ilqr_horizon(const H1Ilqr* h);

/* Replaces RobotUtils::setCostWeights + set*Weight + setConstraintWeights. */
int h1ilqr_set_weights(H1Ilqr* h, const H1Weights* w);

/* Full symmetric cost matrices: RobotUtils::setCostWeights stores whole Q, R, Qf and iLQR multiplies them
 * (ilqr.cpp:145-150: lx = Q (x - x_ref), lxx = Q, lu = R (u - u_ref), luu = R; :372-373, :441: 0.5 e'Qe in the line-search
 * cost). Column-major Q [51 x 51], R [19 x 19], Qf [51 x 51]; they must be symmetric (H1ILQR_EARG otherwise). Their diagonals
 * replace Qdiag / Rdiag / Qfdiag of the last h1ilqr_set_weights, which in turn keeps any off-diagonal parts set here. All
 * three NULL = diagonal weights again. */
int h1ilqr_set_weight_matrices(H1Ilqr* h, const double* Q, const double* R, const double* Qf);

/* Reference window for every instance. Replaces MPC::extractReferenceWindow /
 * RobotUtils::getReferenceWindow (src/ilqr/mpc.cpp:163-166, robot_utils.cpp:422-443) plus the
 * horizon-local lookups isStance / getEEReference / getCoMVelReference (robot_utils.cpp:494-549).
 *   x_ref   [batch][N+1][51]   u_ref [batch][N][19]   com_ref [batch][N+1][3]
 *   ee_ref  [batch][N+1][2][3] (left, right ankle)    stance  [batch][N+1][2] (1 = stance)
 *   com_vel_ref [batch][N+1][3] (may be NULL when w_com_vel == 0)
 * If `shared` != 0 the arrays hold ONE window ([1][...]) used by all instances. */
int h1ilqr_set_reference_window(H1Ilqr* h, const double* x_ref, const double* u_ref, const double* com_ref,
                                const double* ee_ref, const int* stance, const double* com_vel_ref,
                                int shared);

/* Initial guess. Replaces iLQR::initializeWithReference (src/ilqr/ilqr.cpp:50-117).
 *  warm[i] != 0 : shift instance i's previous solution by one knot and roll out the last step;
 *  warm[i] == 0 : cold start, ubar[t] = u_init (19 values per instance, or shared if u_init_shared)
 *                 followed by a full rollout (the reference's gravity-compensation guess, Q15, is an
 *                 explicit input here).
 *  x0 [batch][51]; warm [batch] (NULL = all cold); u_init [batch][19] or [19]. */
int h1ilqr_initialize(H1Ilqr* h, const double* x0, const int* warm, const double* u_init, int u_init_shared);

/* Full multi-iteration solve for all instances. Replaces iLQR::solve (src/ilqr/ilqr.cpp:521-660).
 * Outputs (any may be NULL): cost_out[batch], iters_out[batch], status_out[batch] (0 ok, 1 non-finite).
 * The per-instance regularisation lambda persists across calls, as reg_lambda_ does (ilqr.hpp:54). */
int h1ilqr_solve(H1Ilqr* h, const double* x0, double* cost_out, int* iters_out, int* status_out);

/* MPC step for all instances: initialize (warm where a previous solution exists) + solve +
 * u_apply = ubar[0] + K[0](x_measured - xbar[0]) + store previous solution.
 * Replaces MPC::stepOnce (src/ilqr/mpc.cpp:40-127). u_apply [batch][19]. Returns H1ILQR_ENOTFINITE (with u_apply /
 * cost_out filled in) when an instance produced a non-finite cost or gains; h1ilqr_get_status tells which. */
int h1ilqr_mpc_step(H1Ilqr* h, const double* x_measured, const double* u_init, int u_init_shared,
                    double* u_apply, double* cost_out);
int h1ilqr_mpc_reset(H1Ilqr* h);
/* Per-instance status (0 ok, 1 non-finite cost / gains) and iteration count of the last solve / MPC step.
 * h1ilqr_mpc_step and h1ilqr_solve return H1ILQR_ENOTFINITE (outputs still delivered) when any status is 1. */
int h1ilqr_get_status(H1Ilqr* h, int* status_out, int* iters_out);

/* Kernel families. The path has two sm_100a implementations of its per-knot / per-rollout stages with identical
 * semantics (both are parity-tested against the oracle): COOPERATIVE = one warp per unit (dynamics evaluation,
 * candidate rollout, knot), BATCHED = one thread per unit for the nominal rollout / factorisation / tangent
 * directions and one warp per instance (8 candidates x 4 kinematic chains) for the line search. AUTO (default)
 * chooses per stage by batch size; its line search is the BATCHED one at every batch size (it is also the faster
 * one for a single instance). There is no CPU path in either. */
#define H1ILQR_KERNELS_AUTO 0
#define H1ILQR_KERNELS_COOPERATIVE 1
#define H1ILQR_KERNELS_BATCHED 2
int h1ilqr_set_kernel_policy(H1Ilqr* h, int policy);

/* ---- granular stages (used by the parity tests and by the iLQR shim class) ---- */
/* iLQR::forwardRolloutNominal (ilqr.cpp:119-124): xbar[t+1] = f_D(xbar[t], ubar[t]) from xbar[0]=x0. */
int h1ilqr_rollout_nominal(H1Ilqr* h, const double* x0);
/* iLQR::computeLinearization -> RobotUtils::linearizeDynamicsFD (ilqr.cpp:126-131, robot_utils.cpp:120-160). */
int h1ilqr_linearize(H1Ilqr* h);
/* iLQR::computeCostQuadratics (ilqr.cpp:133-244) incl. derivatives.cpp cost terms and limit penalties. */
int h1ilqr_cost_quadratics(H1Ilqr* h);
/* iLQR::backwardPass (ilqr.cpp:250-309) with the instance's current lambda. */
int h1ilqr_backward_pass(H1Ilqr* h);
/* iLQR::forwardPassLineSearch (ilqr.cpp:311-361): improved[batch], new_cost[batch], alpha_index[batch] (-1 = none). */
int h1ilqr_line_search(H1Ilqr* h, const double* x0, int* improved, double* new_cost, int* alpha_index);
/* iLQR::computeTotalCost (ilqr.cpp:363-518) of the current xbar/ubar. */
int h1ilqr_total_cost(H1Ilqr* h, double* cost_out);

/* One-step dynamics for arbitrary states: RobotUtils::rolloutOneStep (robot_utils.cpp:106-117) and
 * RobotUtils::step (robot_utils.cpp:99-103). x [n][51], u [n][19] -> x_next [n][51]. n <= batch*(N+1). */
int h1ilqr_dynamics_step(H1Ilqr* h, int n, const double* x, const double* u, double* x_next);
/* qfrc_bias analogue (Coriolis + gravity, 25 per state): RobotUtils::computeGravComp (robot_utils.cpp:844-866). */
int h1ilqr_bias_forces(H1Ilqr* h, int n, const double* x, double* bias);
/* Dynamics-model FK used to precompute references: CoM (subtree_com of the root) and ankle body
 * positions, RobotUtils::loadReferences (robot_utils.cpp:370-403). com [n][3], ee [n][2][3]. */
int h1ilqr_reference_kinematics(H1Ilqr* h, int n, const double* x, double* com, double* ee);
/* Whole-body CoM velocity (world frame) of arbitrary states on the dynamics model: the per-row CoM-velocity target
 * J_subtreeCom(root) * qvel of RobotUtils::loadReferences (robot_utils.cpp:388-397), tracked by
 * addCoMVelCostDerivatives when W_com_vel > 0 (ilqr.cpp:675-695). com_vel [n][3]. */
int h1ilqr_reference_com_velocity(H1Ilqr* h, int n, const double* x, double* com_vel);
/* World velocity of the two ankle-body origins: the ee_vel_ref rows jac_pos * qvel of RobotUtils::loadReferences
 * (robot_utils.cpp:405-412; RobotUtils::getEEVelReference). ee_vel [n][2][3]. */
int h1ilqr_reference_ee_velocity(H1Ilqr* h, int n, const double* x, double* ee_vel);
/* One (x, u) pair linearized on its own: RobotUtils::linearizeDynamicsFD (robot_utils.cpp:120-160, include/common/
 * robot_utils.hpp:51-53) for mode == H1ILQR_LIN_FD (forward differences with step eps), the exact Jacobians of f_D for
 * H1ILQR_LIN_ANALYTIC. A [51 x 51], B [51 x 19] column-major. The solver state of the handle is not touched. */
int h1ilqr_linearize_state(H1Ilqr* h, int mode, double eps, const double* x, const double* u, double* A, double* B);
/* RobotUtils::constraintCost / constraintGradients / constraintHessians (robot_utils.cpp:615-778) for n (x, u) pairs
 * (u == NULL: joint-limit terms only, the terminal form of robot_utils.cpp:226-250). Outputs, any may be NULL: cost [n],
 * grad_x [n][51], grad_u [n][19], and the DIAGONALS of the diagonal Hessians, hess_xx_diag [n][51], hess_uu_diag [n][19]. */
int h1ilqr_limit_penalties(H1Ilqr* h, int n, const double* x, const double* u, double* cost, double* grad_x, double* grad_u,
                           double* hess_xx_diag, double* hess_uu_diag);
/* RobotUtils::stageCost (u != NULL) / terminalCost (u == NULL), robot_utils.cpp:162-252: 0.5 e'Qe + 0.5 eu'R eu +
 * 0.5 W_com |com(x) - com_ref|^2 + limit penalties, for n states against the reference rows x_ref [n][51], u_ref [n][19]
 * (NULL = zeros), com_ref [n][3] (NULL = no CoM term), with the handle's diagonal Q / R / Qf. cost [n]. */
int h1ilqr_stage_cost(H1Ilqr* h, int n, const double* x, const double* u, const double* x_ref, const double* u_ref,
                      const double* com_ref, double* cost);
/* World positions of the 2 x 4 sole contact points of f_D for arbitrary states, pts [n][8][3] (left foot first).
 * Input of the contact-schedule generation that replaces get_contacts.py:96-157 (MuJoCo foot-geom contacts with
 * dist < 1e-3): a foot is in stance when one of its sole points is lower than the threshold. */
int h1ilqr_sole_points(H1Ilqr* h, int n, const double* x, double* pts);

/* ---- accessors (host copies). Sizes as in the conventions above. Any pointer may be NULL. ---- */
int h1ilqr_set_trajectory(H1Ilqr* h, const double* xbar, const double* ubar);
int h1ilqr_get_trajectory(H1Ilqr* h, double* xbar, double* ubar);
int h1ilqr_get_gains(H1Ilqr* h, double* K, double* kff);
int h1ilqr_set_gains(H1Ilqr* h, const double* K, const double* kff);
int h1ilqr_get_linearization(H1Ilqr* h, double* A, double* B);
int h1ilqr_set_linearization(H1Ilqr* h, const double* A, const double* B);
int h1ilqr_get_cost_quadratics(H1Ilqr* h, double* lx, double* lu, double* lxx, double* luu);
int h1ilqr_set_cost_quadratics(H1Ilqr* h, const double* lx, const double* lu, const double* lxx, const double* luu);
/* Previous MPC solution that a warm start shifts by one knot: the prev_xbar / prev_ubar arguments of
 * iLQR::initializeWithReference (ilqr.hpp:41-45), owned by MPC (mpc.hpp:57-58). h1ilqr_mpc_step maintains it on the
 * device; a caller that owns the previous solution itself (the iLQR shim class) hands it over here and then calls
 * h1ilqr_initialize with warm = 1. prev_xbar [batch][N+1][51], prev_ubar [batch][N][19]. */
int h1ilqr_set_previous_solution(H1Ilqr* h, const double* prev_xbar, const double* prev_ubar);
int h1ilqr_get_previous_solution(H1Ilqr* h, double* prev_xbar, double* prev_ubar);
int h1ilqr_get_regularization(H1Ilqr* h, double* lambda);
int h1ilqr_set_regularization(H1Ilqr* h, const double* lambda, int shared);
/* per-iteration trace of the last solve: cost_trace [batch][max_iterations], alpha_trace [batch][max_iterations][2] */
int h1ilqr_get_solve_trace(H1Ilqr* h, double* cost_trace, int* alpha_trace);

/* ---- device-resident stepping (measurement + closed-loop batches without host round trips) ----
 * h1ilqr_upload_inputs stages x_measured / u_init once; h1ilqr_run_resident_steps then enqueues `steps` complete
 * MPC steps (initialize + solve + first control, exactly h1ilqr_mpc_step's kernel sequence) on the handle's
 * stream with NO host<->device copies inside, brackets them with CUDA events on that stream and returns the
 * elapsed milliseconds. cold_each_step != 0 forgets the previous solution and resets lambda before every step
 * so that every step is the same cold-start solve. */
int h1ilqr_upload_inputs(H1Ilqr* h, const double* x_measured, const double* u_init, int u_init_shared);
/* cold_each_step: bit 0 = cold start at every step; bit 1 = capture the step once into a CUDA graph and replay it (the
 * launch sequence of a step is fixed: iteration counts are handled on the device by the compact instance lists). */
int h1ilqr_run_resident_steps(H1Ilqr* h, int steps, int cold_each_step, double* elapsed_ms);

/* ---- device-resident closed loop: main/humanoid_mpc.cpp:130-179 for the whole batch without a host round
 * trip per step: getState -> MPC::stepOnce (window at t_idx, warm start, solve, first control) -> setControl -> step.
 * h1ilqr_set_reference_table uploads RobotUtils' full tables once (x_ref_full_ [rows][51], com_ref_full_ [rows][3],
 * ee_pos_ref_full_ [rows][2][3], com_vel_ref_full_ [rows][3] (NULL = zeros), contact_schedule_ [contact_rows][2]; loadReferences
 * / loadContactSchedule, robot_utils.cpp:281-492). Windows are extracted on the device with the reference's rules: rows
 * min(t_idx + i, last) for x_ref / com_ref (robot_utils.cpp:422-443), HORIZON-LOCAL rows i for isStance / getEEReference /
 * getCoMVelReference (quirk Q6; schedule_offset != 0 uses rows t_idx + i instead), u_ref = 0.
 * h1ilqr_run_closed_loop runs `steps` closed-loop steps of every instance: plant = f_D (RobotUtils::step, robot_utils.cpp:
 * 99-103), time index t_idx0[i] + step (NULL: continue from the current indices, which start at 0), warm start from the
 * previous solution kept on the device (h1ilqr_mpc_reset forgets it), lambda persistent. x_start NULL = continue from the
 * current plant states. Outputs (any may be NULL): x_final [batch][51], cost_log / iters_log [steps][batch],
 * u_log [steps][batch][19] (applied controls), elapsed_ms (CUDA events around the whole loop). use_graph: replay one captured
 * step. Returns H1ILQR_ENOTFINITE when an instance is non-finite at the last step. */
int h1ilqr_set_reference_table(H1Ilqr* h, int rows, const double* x_ref_full, const double* com_ref_full,
                               const double* ee_ref_full, const double* com_vel_ref_full, int contact_rows,
                               const int* contact, int schedule_offset);
int h1ilqr_run_closed_loop(H1Ilqr* h, int steps, const int* t_idx0, const double* x_start, const double* u_init,
                           int u_init_shared, int use_graph, double* x_final, double* cost_log, int* iters_log,
                           double* u_log, double* elapsed_ms);
/* Device time (CUDA events on the handle's stream, milliseconds summed over `reps` launches) of one stage on the handle's
 * current trajectory / derivatives / gains: 0 factorisation of Mhat, 1 linearization, 2 cost quadratics, 3 backward
 * pass, 4 line search (the trajectory is restored after every repetition). Measurement only. */
int h1ilqr_time_stage(H1Ilqr* h, int stage, int reps, double* elapsed_ms);
/* sustained fp64 FMA throughput of the device (TFLOP/s), measured with a register-resident DFMA kernel: the
 * FP64 roofline denominator (MEASURED_PEAKS.json carries none). */
int h1ilqr_measure_fp64_peak(H1Ilqr* h, double* tflops);
/* same for the fp64 tensor-core path (mma.sync m8n8k4, SASS DMMA) used by the Riccati contractions. */
int h1ilqr_measure_fp64_mma_peak(H1Ilqr* h, double* tflops);

/* Page-lock (pin) a caller-owned host buffer so that the host <-> device copies of h1ilqr_set_reference_window /
 * h1ilqr_mpc_step / the getters run as direct DMA transfers instead of going through the driver's bounce buffers.
 * The MPC loop keeps its reference arrays alive across steps (MPC members x_ref_window_ ..., src/ilqr/mpc.hpp:52-60),
 * so they are registered once. Buffers that are not registered keep working (pageable copies). */
int h1ilqr_host_register(H1Ilqr* h, const void* host_ptr, size_t bytes);
int h1ilqr_host_unregister(H1Ilqr* h, const void* host_ptr);

/* ---- timing of the last h1ilqr_solve, CUDA events on the handle's stream, milliseconds ---- */
typedef struct H1StageTimes {
  double total_ms;
  double rollout_ms, linearize_ms, cost_quadratics_ms, backward_ms, line_search_ms;
  int launches; /* kernels launched by the last solve */
} H1StageTimes;
int h1ilqr_enable_stage_timing(H1Ilqr* h, int enable);
int h1ilqr_get_stage_times(H1Ilqr* h, H1StageTimes* t);
/* raw CUDA stream (cudaStream_t) the handle launches on, for external event timing */
void* h1ilqr_stream(H1Ilqr* h);

#ifdef __cplusplus
}
#endif
#endif
