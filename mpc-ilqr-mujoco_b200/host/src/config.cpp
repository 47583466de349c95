// loadConfigFromFile + Config::buildCostMatrices (reference: src/common/config.cpp:4-122) with a small reader
// for the YAML subset config.yaml uses: nested maps by indentation, scalars, inline [a, b, c] lists, quoted
// strings, '#' comments. Unknown keys are ignored, missing keys are fatal (the reference exits on a YAML error).
#include "common/config.hpp"
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace {

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string strip_comment(const std::string& s) {
  bool in_q = false;
  char q = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    if (in_q) { if (s[i] == q) in_q = false; }
    else if (s[i] == '"' || s[i] == '\'') { in_q = true; q = s[i]; }
    else if (s[i] == '#') return s.substr(0, i);
  }
  return s;
}
std::string unquote(const std::string& s) {
  if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) return s.substr(1, s.size() - 2);
  return s;
}

// flat map "a.b.c" -> raw scalar text
std::map<std::string, std::string> parse_yaml_subset(const std::string& path) {
  std::ifstream f(path);
  if (!f.is_open()) throw std::runtime_error("bad file: " + path);
  std::map<std::string, std::string> out;
  std::vector<std::pair<int, std::string>> stack;  // (indent, key)
  std::string line;
  while (std::getline(f, line)) {
    line = strip_comment(line);
    if (trim(line).empty()) continue;
    int indent = static_cast<int>(line.find_first_not_of(' '));
    std::string body = trim(line);
    size_t colon = body.find(':');
    if (colon == std::string::npos) throw std::runtime_error("cannot parse line: " + body);
    std::string key = trim(body.substr(0, colon)), val = trim(body.substr(colon + 1));
    while (!stack.empty() && stack.back().first >= indent) stack.pop_back();
    std::string full;
    for (auto& s : stack) full += s.second + ".";
    full += key;
    if (val.empty()) stack.push_back({indent, key});
    else out[full] = val;
  }
  return out;
}

struct Yaml {
  std::map<std::string, std::string> kv;
  const std::string& raw(const std::string& k) const {
    auto it = kv.find(k);
    if (it == kv.end()) throw std::runtime_error("missing key: " + k);
    return it->second;
  }
  std::string str(const std::string& k) const { return unquote(raw(k)); }
  double num(const std::string& k) const {
    size_t pos = 0;
    const std::string& r = raw(k);
    double v = std::stod(r, &pos);
    if (trim(r.substr(pos)).size()) throw std::runtime_error("bad conversion: " + k);
    return v;
  }
  int integer(const std::string& k) const { return static_cast<int>(num(k)); }
  bool boolean(const std::string& k) const {
    std::string r = str(k);
    if (r == "true" || r == "True" || r == "yes" || r == "on") return true;
    if (r == "false" || r == "False" || r == "no" || r == "off") return false;
    throw std::runtime_error("bad conversion: " + k);
  }
  std::vector<double> list(const std::string& k) const {
    std::string r = raw(k);
    if (r.size() < 2 || r.front() != '[' || r.back() != ']') throw std::runtime_error("bad conversion: " + k);
    std::vector<double> v;
    std::stringstream ss(r.substr(1, r.size() - 2));
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(trim(tok)));
    return v;
  }
};

}  // namespace

Config loadConfigFromFile(const std::string& filepath) {
  Config config;
  try {
    Yaml y{parse_yaml_subset(filepath)};
    config.model_path = y.str("robot.model_path");
    config.urdf_path = y.str("robot.urdf_path");
    config.q_ref_path = y.str("reference_trajectory.q_ref");
    config.v_ref_path = y.str("reference_trajectory.v_ref");
    config.contact_schedule_path = y.str("reference_trajectory.contact_schedule");
    config.results_path = y.str("logging.results_path");
    config.verbose = y.boolean("logging.verbose");
    config.save_trajectories = y.boolean("logging.save_trajectories");
    config.mpc.horizon = y.integer("mpc.horizon");
    config.mpc.dt = y.num("mpc.dt");
    config.mpc.physics_dt = y.num("mpc.physics_dt");
    config.mpc.gravity = y.list("mpc.gravity");
    config.mpc.sim_steps = y.integer("mpc.sim_steps");
    config.mpc.contact_impratio = y.num("mpc.contact_impratio");
    CostWeights& c = config.mpc.costs;
    const std::string p = "mpc.cost_weights.";
    c.Q_position_x = y.num(p + "Q_position_x"); c.Q_position_y = y.num(p + "Q_position_y");
    c.Q_position_z = y.num(p + "Q_position_z"); c.Q_quat_w = y.num(p + "Q_quat_w");
    c.Q_quat_xyz = y.list(p + "Q_quat_xyz"); c.Q_joint_pos = y.num(p + "Q_joint_pos");
    c.Q_vel_x = y.num(p + "Q_vel_x"); c.Q_vel_y = y.num(p + "Q_vel_y"); c.Q_vel_z = y.num(p + "Q_vel_z");
    c.Q_ang_vel = y.num(p + "Q_ang_vel"); c.Q_joint_vel = y.num(p + "Q_joint_vel");
    c.R_control = y.num(p + "R_control"); c.Qf_multiplier = y.num(p + "Qf_multiplier");
    c.Qf_position_x = y.num(p + "Qf_position_x"); c.Qf_position_y = y.num(p + "Qf_position_y");
    c.Qf_position_z = y.num(p + "Qf_position_z"); c.Qf_vel_z = y.num(p + "Qf_vel_z");
    c.W_com = y.num(p + "W_com_pos"); c.W_com_vel = y.num(p + "W_com_vel");
    c.W_foot = y.num(p + "W_foot"); c.W_foot_vel = y.num(p + "W_foot_vel");
    c.W_upright = y.num(p + "W_upright"); c.w_balance = y.num(p + "w_balance");
    config.mpc.joint_limit_weight = y.num("mpc.constraints.joint_limit_weight");
    config.mpc.torque_limit_weight = y.num("mpc.constraints.torque_limit_weight");
    if (c.Q_quat_xyz.size() != 3 || config.mpc.gravity.size() != 3) throw std::runtime_error("bad list length");
  } catch (const std::exception& e) {
    std::cerr << "Failed to load or parse config.yaml: " << e.what() << std::endl;
    std::exit(1);
  }
  return config;
}

void Config::buildCostMatrices(int nx, int nu, int nq) {
  Q = Eigen::MatrixXd::Identity(nx, nx);
  R = Eigen::MatrixXd::Identity(nu, nu);
  const CostWeights& c = mpc.costs;
  Q(0, 0) = c.Q_position_x; Q(1, 1) = c.Q_position_y; Q(2, 2) = c.Q_position_z;
  Q(3, 3) = c.Q_quat_w; Q(4, 4) = c.Q_quat_xyz[0]; Q(5, 5) = c.Q_quat_xyz[1]; Q(6, 6) = c.Q_quat_xyz[2];
  for (int i = 7; i < nq; ++i) Q(i, i) = c.Q_joint_pos;
  Q(nq + 0, nq + 0) = c.Q_vel_x; Q(nq + 1, nq + 1) = c.Q_vel_y; Q(nq + 2, nq + 2) = c.Q_vel_z;
  for (int i = 3; i < 6; ++i) Q(nq + i, nq + i) = c.Q_ang_vel;
  for (int i = nq + 6; i < nx; ++i) Q(i, i) = c.Q_joint_vel;
  R *= c.R_control;
  Qf = Q * c.Qf_multiplier;
  Qf(0, 0) *= c.Qf_position_x; Qf(1, 1) *= c.Qf_position_y; Qf(2, 2) *= c.Qf_position_z;
  Qf(nq + 2, nq + 2) *= c.Qf_vel_z;
  std::cout << "Cost matrices built: Q(" << nx << "x" << nx << "), R(" << nu << "x" << nu << "), Qf(" << nx << "x" << nx
            << ")" << std::endl;
}
