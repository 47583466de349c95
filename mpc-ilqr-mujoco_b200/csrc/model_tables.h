#pragma once
#include "h1_common.cuh"

namespace h1 {
bool build_dyn_model(const H1Model& m, DynModel* d);
bool build_cost_model(const H1Model& m, CostModel* c);
}  // namespace h1
