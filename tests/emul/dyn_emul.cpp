// Lane-emulation build of the warp-cooperative f_D (mpc-ilqr-mujoco_b200/csrc/h1_dyn.cuh) for the CPU test
// suite: the phase functions are compiled as plain C++ and the 32 lanes of a phase run one after another.
// This checks the kernel's index tables / sparse factorisation logic without a GPU; it is not a product path.
#include "../../mpc-ilqr-mujoco_b200/csrc/h1_dyn.cuh"
#include "../../mpc-ilqr-mujoco_b200/csrc/h1_lin_dirs.cuh"
#include "../../mpc-ilqr-mujoco_b200/csrc/h1_dyn_seq.cuh"
#include "../../mpc-ilqr-mujoco_b200/csrc/model_tables.h"
#include <cstring>

extern "C" int emul_dyn_step(int n, const double* x, const double* u, double* xn, double* com) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  static h1::DynWarp w;
  for (int i = 0; i < n; ++i) {
    h1::dyn_step_warp(md, w, x + i * h1::NX, u + i * h1::NU, xn + i * h1::NX);
    if (com) for (int k = 0; k < 3; ++k) com[3 * i + k] = w.com[k];
  }
  return 0;
}

// exact linearization through the tangent phases (csrc/h1_dyn.cuh), column-major A[51x51], B[51x19]
extern "C" int emul_dyn_linearize_analytic(const double* x, const double* u, double* A, double* B) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  static h1::DynWarp w;
  static h1::DynWarpT<h1::Dual> wd;
  static h1::PrimalFactor pf;
  h1::dyn_primal_factor_warp(md, w, x, u, nullptr, pf);
  for (int e = 0; e < h1::NX + h1::NU; ++e) {
    h1::dyn_tangent_assemble_warp(md, wd, x, u, e);
    h1::dyn_tangent_solve_warp(md, wd, pf, e < h1::NX ? A + e * h1::NX : B + (e - h1::NX) * h1::NX);
  }
  return 0;
}

// same, through the inverse-dynamics tangent (the path the kernel k_linearize_analytic runs)
extern "C" int emul_dyn_linearize_id(const double* x, const double* u, double* A, double* B) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  static h1::DynWarp w;
  static h1::TanWarpT<h1::Dual> wt;
  static h1::PrimalFactor pf;
  h1::dyn_primal_factor_warp(md, w, x, u, nullptr, pf);
  for (int e = 0; e < h1::NX + h1::NU; ++e)
    h1::dyn_tangent_id_warp(md, wt, pf, x, u, e, e < h1::NX ? A + e * h1::NX : B + (e - h1::NX) * h1::NX);
  return 0;
}

// same columns through the direction-per-thread functions (csrc/h1_lin_dirs.cuh, kernel k_linearize_dirs)
extern "C" int emul_dyn_linearize_dirs(const double* x, const double* u, double* A, double* B) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  static h1::DynWarp w;
  static h1::PrimalFactor pf;
  h1::dyn_primal_factor_warp(md, w, x, u, nullptr, pf);
  for (int e = 0; e < h1::NX + h1::NU; ++e) {
    double tv[h1::NV];
    if (e < h1::NQ) h1::id_tangent_seq<h1::Dual, h1::Dual>(md, x, pf.a, e, tv);
    else if (e < h1::NX) h1::id_tangent_seq<double, h1::Dual>(md, x, pf.a, e, tv);
    else {
      const int j = e - h1::NX;
      for (int k = 0; k < h1::NV; ++k) tv[k] = 0.0;
      tv[6 + j] = (u[j] < md.ctrl_lo[j] || u[j] > md.ctrl_hi[j]) ? 0.0 : 1.0;
    }
    if (md.seq_ok && (e & 1)) h1::tangent_solve_h1(&pf.Lm[0][0], pf.D, tv);   // both solve variants are exercised
    else h1::tangent_solve_seq(md, &pf.Lm[0][0], pf.D, tv);
    h1::integrate_tangent_seq(md, x, pf.a, e, tv, e < h1::NX ? A + e * h1::NX : B + (e - h1::NX) * h1::NX);
  }
  return 0;
}

// thread-sequential f_D (csrc/h1_dyn_seq.cuh, kernels k_rollout_seq / k_line_search_seq); also returns the factor
// it would hand to the linearization kernels next to the warp-cooperative one, for comparison
extern "C" int emul_dyn_step_seq(int n, const double* x, const double* u, double* xn, double* com, double* fac_seq,
                                 double* fac_warp) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  static h1::DynWarp w;
  const int nf = sizeof(h1::PrimalFactor) / sizeof(double);
  for (int i = 0; i < n; ++i) {
    h1::PrimalFactor pf, pw;
    std::memset(&pf, 0, sizeof(pf)); std::memset(&pw, 0, sizeof(pw));
    h1::dyn_step_seq(md, x + i * h1::NX, u + i * h1::NU, xn + i * h1::NX, &pf, com + 3 * i);
    h1::dyn_primal_factor_warp(md, w, x + i * h1::NX, u + i * h1::NU, nullptr, pw);
    std::memcpy(fac_seq + (size_t)i * nf, &pf, sizeof(pf));
    std::memcpy(fac_warp + (size_t)i * nf, &pw, sizeof(pw));
  }
  return 0;
}
extern "C" int emul_dyn_com_seq(int n, const double* x, double* com) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  for (int i = 0; i < n; ++i) h1::dyn_com_seq(md, x + i * h1::NX, com + 3 * i);
  return 0;
}

// same columns through the per-direction tangent functions of kernel k_linearize_tangents (csrc/h1_lin_dirs.cuh): joint
// directions walk only subtree(joint) with dual numbers, x / y columns are unit vectors
extern "C" int emul_dyn_linearize_cols(const double* x, const double* u, double* A, double* B) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  static h1::DynWarp w;
  static h1::PrimalFactor pf;
  h1::dyn_primal_factor_warp(md, w, x, u, nullptr, pf);
  for (int e = 0; e < h1::NX + h1::NU; ++e) {
    double* col = e < h1::NX ? A + e * h1::NX : B + (e - h1::NX) * h1::NX;
    h1::linearize_column(md, x, u, pf, e, col);
  }
  return 0;
}
