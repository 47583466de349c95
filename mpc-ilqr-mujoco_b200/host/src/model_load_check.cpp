// Parses an MJCF and a URDF with the run-time loaders and prints both H1Model tables as JSON (no GPU needed):
// tests/test_host.py compares them with the tables generated at build time. Usage: model_load_check scene_or_h1.xml h1.urdf
#include <iostream>
#include "common/model_loader.hpp"

static void arr(const char* name, const double* v, int n, bool last = false) {
  std::cout << "\"" << name << "\": [";
  for (int i = 0; i < n; ++i) std::cout << (i ? ", " : "") << v[i];
  std::cout << "]" << (last ? "" : ", ");
}
static void iarr(const char* name, const int* v, int n) {
  std::cout << "\"" << name << "\": [";
  for (int i = 0; i < n; ++i) std::cout << (i ? ", " : "") << v[i];
  std::cout << "], ";
}
static void dump(const H1Model& m) {
  std::cout << "{";
  iarr("parent", m.parent, H1_NB); iarr("axis", m.axis, H1_NB); iarr("has_rfix", m.has_rfix, H1_NB); iarr("foot_body", m.foot_body, 2);
  arr("pos", &m.pos[0][0], 3 * H1_NB); arr("rfix", &m.rfix[0][0], 9 * H1_NB); arr("mass", m.mass, H1_NB);
  arr("ipos", &m.ipos[0][0], 3 * H1_NB); arr("inertia", &m.inertia[0][0], 6 * H1_NB); arr("armature", m.armature, H1_NV);
  arr("damping", m.damping, H1_NV); arr("jnt_range", &m.jnt_range[0][0], 2 * H1_NU); arr("ctrl_range", &m.ctrl_range[0][0], 2 * H1_NU);
  arr("foot_pts", &m.foot_pts[0][0][0], 24); arr("gravity", m.gravity, 3);
  const double sc[6] = {m.timestep, m.contact_kn, m.contact_bn, m.contact_bt, m.contact_eps, m.total_mass};
  arr("scalars", sc, 6, true);
  std::cout << "}";
}

int main(int argc, char** argv) {
  if (argc < 3) { std::cerr << "usage: model_load_check model.xml model.urdf" << std::endl; return 2; }
  H1Model dyn, cost;
  std::vector<std::string> jn, bn;
  std::string err;
  if (!load_mjcf_model(argv[1], *h1_default_dynamics_model(), &dyn, &jn, &bn, &err)) { std::cerr << "MJCF: " << err << std::endl; return 1; }
  if (!load_urdf_model(argv[2], *h1_default_cost_model(), jn, bn, &cost, &err)) { std::cerr << "URDF: " << err << std::endl; return 1; }
  std::cout.precision(17);
  std::cout << "{\"joints\": [";
  for (size_t i = 0; i < jn.size(); ++i) std::cout << (i ? ", " : "") << "\"" << jn[i] << "\"";
  std::cout << "], \"dynamics\": "; dump(dyn);
  std::cout << ", \"cost\": "; dump(cost);
  std::cout << "}" << std::endl;
  return 0;
}
