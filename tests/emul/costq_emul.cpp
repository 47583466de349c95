// Lane-emulation build of the warp-cooperative cost-quadratics kernel (csrc/h1_costq.cuh) for CPU tests.
#include "../../mpc-ilqr-mujoco_b200/csrc/h1_costq.cuh"
#include "../../mpc-ilqr-mujoco_b200/csrc/model_tables.h"

extern "C" int emul_cost_quadratics(const H1Weights* wt, const double* x, const double* u, const double* x_ref,
                                    const double* u_ref, const double* com_ref, const double* com_vel_ref,
                                    const double* ee_ref, const int* stance, int terminal, double* lx, double* lu,
                                    double* lxx, double* luu) {
  static h1::DynModel md;
  static h1::CostModel cm;
  static bool init = false;
  if (!init) {
    if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1;
    if (!h1::build_cost_model(*h1_default_cost_model(), &cm)) return -1;
    init = true;
  }
  static h1::CostWarp w;
  h1::KnotTargets kt{com_ref, com_vel_ref, ee_ref, stance, terminal != 0};
  h1::cost_quadratics_warp(cm, md, *wt, w, x, u, x_ref, u_ref, kt, lx, lu, lxx, luu);
  return 0;
}
