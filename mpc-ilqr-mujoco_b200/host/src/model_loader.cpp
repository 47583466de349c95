// Run-time MJCF / URDF readers for the H1 model class (see common/model_loader.hpp).
#include "common/model_loader.hpp"
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace {

// ---------------- a minimal XML element tree: tags, double- or single-quoted attributes, comments, declarations ----------------
struct Xml {
  std::string tag;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<Xml>> kids;
  const Xml* child(const std::string& t) const { for (auto& k : kids) if (k->tag == t) return k.get(); return nullptr; }
  std::vector<const Xml*> children(const std::string& t) const { std::vector<const Xml*> v; for (auto& k : kids) if (k->tag == t) v.push_back(k.get()); return v; }
  bool has(const std::string& a) const { return attr.count(a) != 0; }
  std::string get(const std::string& a, const std::string& dflt = "") const { auto it = attr.find(a); return it == attr.end() ? dflt : it->second; }
};

struct XmlParser {
  const std::string& s;
  size_t i = 0;
  explicit XmlParser(const std::string& text) : s(text) {}
  void skip_ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  bool starts(const char* p) const { return s.compare(i, std::strlen(p), p) == 0; }
  void skip_misc() {   // whitespace, comments, <?...?>, <!DOCTYPE ...>, text
    for (;;) {
      while (i < s.size() && s[i] != '<') ++i;
      if (i >= s.size()) return;
      if (starts("<!--")) { size_t e = s.find("-->", i); if (e == std::string::npos) throw std::runtime_error("unterminated comment"); i = e + 3; }
      else if (starts("<?")) { size_t e = s.find("?>", i); if (e == std::string::npos) throw std::runtime_error("unterminated declaration"); i = e + 2; }
      else if (starts("<!")) { size_t e = s.find('>', i); if (e == std::string::npos) throw std::runtime_error("unterminated <!"); i = e + 1; }
      else return;
    }
  }
  std::string name() { size_t b = i; while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.')) ++i; return s.substr(b, i - b); }
  std::unique_ptr<Xml> element() {
    skip_misc();
    if (i >= s.size() || s[i] != '<' || starts("</")) return nullptr;
    ++i;
    std::unique_ptr<Xml> e(new Xml);
    e->tag = name();
    if (e->tag.empty()) throw std::runtime_error("bad element name");
    for (;;) {
      skip_ws();
      if (i >= s.size()) throw std::runtime_error("unterminated element <" + e->tag + ">");
      if (starts("/>")) { i += 2; return e; }
      if (s[i] == '>') { ++i; break; }
      std::string a = name();
      skip_ws();
      if (a.empty() || i >= s.size() || s[i] != '=') throw std::runtime_error("bad attribute in <" + e->tag + ">");
      ++i; skip_ws();
      const char q = s[i];
      if (q != '"' && q != '\'') throw std::runtime_error("unquoted attribute in <" + e->tag + ">");
      size_t end = s.find(q, i + 1);
      if (end == std::string::npos) throw std::runtime_error("unterminated attribute value");
      e->attr[a] = s.substr(i + 1, end - i - 1);
      i = end + 1;
    }
    for (;;) {
      skip_misc();
      if (i >= s.size()) throw std::runtime_error("missing </" + e->tag + ">");
      if (starts("</")) {
        i += 2;
        if (name() != e->tag) throw std::runtime_error("mismatched </" + e->tag + ">");
        skip_ws();
        if (i < s.size() && s[i] == '>') ++i;
        return e;
      }
      auto k = element();
      if (k) e->kids.push_back(std::move(k));
    }
  }
};

std::unique_ptr<Xml> parse_file(const std::string& path) {
  std::ifstream f(path);
  if (!f.is_open()) throw std::runtime_error("cannot open " + path);
  std::stringstream ss; ss << f.rdbuf();
  const std::string text = ss.str();
  XmlParser p(text);
  auto root = p.element();
  if (!root) throw std::runtime_error("no root element in " + path);
  return root;
}

std::vector<double> numbers(const std::string& str, size_t n, const char* what) {
  std::stringstream ss(str);
  std::vector<double> v; double x;
  while (ss >> x) v.push_back(x);
  if (v.size() != n) throw std::runtime_error(std::string("expected ") + std::to_string(n) + " numbers in " + what + ": '" + str + "'");
  return v;
}
void quat2mat(const std::vector<double>& q, double* R) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  const double M[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                       2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                       2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
  std::memcpy(R, M, sizeof(M));
}
void matmul3(const double* A, const double* B, double* C) {
  double o[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  std::memcpy(C, o, sizeof(o));
}
void rpy2mat(const std::vector<double>& rpy, double* R) {   // Rz(yaw) Ry(pitch) Rx(roll)
  const double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]), cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  const double Rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr}, Ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp}, Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
  double T[9];
  matmul3(Rz, Ry, T); matmul3(T, Rx, R);
}
int axis_index(const std::vector<double>& a) {
  int idx = 0;
  for (int i = 1; i < 3; ++i) if (std::fabs(a[i]) > std::fabs(a[idx])) idx = i;
  for (int i = 0; i < 3; ++i) if (std::fabs(a[i] - (i == idx ? 1.0 : 0.0)) > 1e-12) throw std::runtime_error("hinge axis is not a positive coordinate axis of the body frame");
  return idx;
}
void set_identity(double* R) { const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(R, I, sizeof(I)); }
void sym_to6(const double* I, double* o) { o[0] = I[0]; o[1] = I[4]; o[2] = I[8]; o[3] = I[1]; o[4] = I[2]; o[5] = I[5]; }
void carry_defaults(const H1Model& d, H1Model* m) {
  std::memcpy(m->foot_pts, d.foot_pts, sizeof(d.foot_pts));
  std::memcpy(m->gravity, d.gravity, sizeof(d.gravity));
  m->timestep = d.timestep; m->contact_kn = d.contact_kn; m->contact_bn = d.contact_bn; m->contact_bt = d.contact_bt; m->contact_eps = d.contact_eps;
}

}  // namespace

bool load_mjcf_model(const std::string& path, const H1Model& defaults, H1Model* out, std::vector<std::string>* joint_names,
                     std::vector<std::string>* body_names, std::string* error) {
  try {
    std::unique_ptr<Xml> root = parse_file(path);
    // scene files only <include> the robot (scene.xml -> h1.xml): follow the first include that has a worldbody with a body
    if (!root->child("worldbody") || !root->child("worldbody")->child("body")) {
      const std::string dir = path.find_last_of('/') == std::string::npos ? "" : path.substr(0, path.find_last_of('/') + 1);
      bool found = false;
      for (const Xml* inc : root->children("include")) {
        std::unique_ptr<Xml> r2 = parse_file(dir + inc->get("file"));
        if (r2->child("worldbody") && r2->child("worldbody")->child("body")) { root = std::move(r2); found = true; break; }
      }
      if (!found) throw std::runtime_error("no robot body in " + path);
    }
    double d_damp = 0.0, d_arm = 0.0;   // joint defaults of the robot's default class (h1.xml:4-8)
    if (const Xml* d0 = root->child("default"))
      for (const Xml* d1 : d0->children("default"))
        if (const Xml* j = d1->child("joint")) { d_damp = std::stod(j->get("damping", "0")); d_arm = std::stod(j->get("armature", "0")); break; }
    H1Model m;
    std::memset(&m, 0, sizeof(m));
    carry_defaults(defaults, &m);
    std::vector<std::string> jn, bn;
    int nb = 0;
    // depth-first pre-order with children in document order = MuJoCo's body / joint order
    struct Rec { static void visit(const Xml* e, int parent, H1Model& m, int& nb, std::vector<std::string>& jn, std::vector<std::string>& bn, double damp, double arm) {
      if (nb >= H1_NB) throw std::runtime_error("more than 20 bodies");
      const int b = nb++;
      bn.push_back(e->get("name"));
      const Xml* in = e->child("inertial");
      if (!in) throw std::runtime_error("body without <inertial>: " + e->get("name"));
      m.parent[b] = parent;
      auto pos = numbers(e->get("pos", "0 0 0"), 3, "body pos");
      for (int i = 0; i < 3; ++i) m.pos[b][i] = (b == 0) ? 0.0 : pos[i];     // the base placement is qpos[0:7]
      if (e->has("quat")) { quat2mat(numbers(e->get("quat"), 4, "body quat"), m.rfix[b]); m.has_rfix[b] = 1; }
      else { set_identity(m.rfix[b]); m.has_rfix[b] = 0; }
      m.mass[b] = std::stod(in->get("mass"));
      auto ip = numbers(in->get("pos", "0 0 0"), 3, "inertial pos");
      for (int i = 0; i < 3; ++i) m.ipos[b][i] = ip[i];
      double Ri[9], I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (in->has("quat")) quat2mat(numbers(in->get("quat"), 4, "inertial quat"), Ri); else set_identity(Ri);
      if (in->has("diaginertia")) {
        auto d = numbers(in->get("diaginertia"), 3, "diaginertia");
        double RD[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) RD[3 * i + j] = Ri[3 * i + j] * d[j];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) I[3 * i + j] = RD[3 * i] * Ri[3 * j] + RD[3 * i + 1] * Ri[3 * j + 1] + RD[3 * i + 2] * Ri[3 * j + 2];
      } else if (in->has("fullinertia")) {
        auto f = numbers(in->get("fullinertia"), 6, "fullinertia");   // xx yy zz xy xz yz
        const double F[9] = {f[0], f[3], f[4], f[3], f[1], f[5], f[4], f[5], f[2]};
        std::memcpy(I, F, sizeof(F));
      } else throw std::runtime_error("inertial without diaginertia / fullinertia");
      sym_to6(I, m.inertia[b]);
      const Xml* j = e->child("joint");
      if (b == 0) {
        if (!e->child("freejoint") && !(j && j->get("type") == "free")) throw std::runtime_error("root body has no free joint");
        m.axis[b] = -1;
      } else {
        if (!j) throw std::runtime_error("body without a hinge: " + e->get("name"));
        jn.push_back(j->get("name"));
        m.axis[b] = axis_index(numbers(j->get("axis", "0 0 1"), 3, "joint axis"));
        auto rg = numbers(j->get("range", "0 0"), 2, "joint range");
        m.jnt_range[b - 1][0] = rg[0]; m.jnt_range[b - 1][1] = rg[1];
        m.damping[5 + b] = j->has("damping") ? std::stod(j->get("damping")) : damp;
        m.armature[5 + b] = j->has("armature") ? std::stod(j->get("armature")) : arm;
      }
      for (const Xml* ch : e->children("body")) visit(ch, b, m, nb, jn, bn, damp, arm);
    } };
    Rec::visit(root->child("worldbody")->child("body"), -1, m, nb, jn, bn, d_damp, d_arm);
    if (nb != H1_NB) throw std::runtime_error("expected 20 bodies, found " + std::to_string(nb));
    // actuators: motor i must drive hinge i (u[i] drives dof 6 + i)
    const Xml* act = root->child("actuator");
    if (!act) throw std::runtime_error("no <actuator> section");
    auto motors = act->children("motor");
    if ((int)motors.size() != H1_NU) throw std::runtime_error("expected 19 motors");
    for (int i = 0; i < H1_NU; ++i) {
      if (motors[i]->get("joint") != jn[i]) throw std::runtime_error("actuator order differs from the joint order at " + jn[i]);
      auto cr = numbers(motors[i]->get("ctrlrange"), 2, "ctrlrange");
      m.ctrl_range[i][0] = cr[0]; m.ctrl_range[i][1] = cr[1];
    }
    m.foot_body[0] = m.foot_body[1] = -1;
    for (int b = 0; b < H1_NB; ++b) {
      if (bn[b] == "left_ankle_link") m.foot_body[0] = b;
      if (bn[b] == "right_ankle_link") m.foot_body[1] = b;
    }
    if (m.foot_body[0] < 0 || m.foot_body[1] < 0) throw std::runtime_error("ankle bodies left_ankle_link / right_ankle_link not found");
    m.total_mass = 0.0;
    for (int b = 0; b < H1_NB; ++b) m.total_mass += m.mass[b];
    *out = m;
    if (joint_names) *joint_names = jn;
    if (body_names) *body_names = bn;
    return true;
  } catch (const std::exception& e) {
    if (error) *error = e.what();
    return false;
  }
}

bool load_urdf_model(const std::string& path, const H1Model& defaults, const std::vector<std::string>& joint_names,
                     const std::vector<std::string>& body_names, H1Model* out, std::string* error) {
  try {
    std::unique_ptr<Xml> root = parse_file(path);
    if ((int)joint_names.size() != H1_NU || (int)body_names.size() != H1_NB) throw std::runtime_error("joint / body name lists of the dynamics model are missing");
    std::map<std::string, const Xml*> links, joints;
    for (const Xml* l : root->children("link")) links[l->get("name")] = l;
    for (const Xml* j : root->children("joint")) joints[j->get("name")] = j;
    // links behind fixed joints must be massless: Pinocchio would otherwise merge their inertia into the parent
    for (auto& kv : joints)
      if (kv.second->get("type") == "fixed") {
        const Xml* ch = kv.second->child("child");
        if (ch && links.count(ch->get("link")) && links[ch->get("link")]->child("inertial")) throw std::runtime_error("fixed-joint child with mass: " + ch->get("link"));
      }
    H1Model m;
    std::memset(&m, 0, sizeof(m));
    carry_defaults(defaults, &m);
    auto inertial = [&](const std::string& link, int b) {
      if (!links.count(link) || !links[link]->child("inertial")) throw std::runtime_error("link without <inertial>: " + link);
      const Xml* it = links[link]->child("inertial");
      const Xml* o = it->child("origin");
      auto rpy = numbers(o ? o->get("rpy", "0 0 0") : "0 0 0", 3, "inertial rpy");
      if (rpy[0] != 0.0 || rpy[1] != 0.0 || rpy[2] != 0.0) throw std::runtime_error("rotated inertial frame in " + link);
      auto xyz = numbers(o ? o->get("xyz", "0 0 0") : "0 0 0", 3, "inertial xyz");
      for (int i = 0; i < 3; ++i) m.ipos[b][i] = xyz[i];
      m.mass[b] = std::stod(it->child("mass")->get("value"));
      const Xml* in = it->child("inertia");
      auto g = [&](const char* k) { return std::stod(in->get(k)); };
      m.inertia[b][0] = g("ixx"); m.inertia[b][1] = g("iyy"); m.inertia[b][2] = g("izz");
      m.inertia[b][3] = g("ixy"); m.inertia[b][4] = g("ixz"); m.inertia[b][5] = g("iyz");
    };
    m.parent[0] = -1; m.axis[0] = -1; m.has_rfix[0] = 0; set_identity(m.rfix[0]);
    inertial(body_names[0], 0);
    for (int i = 0; i < H1_NU; ++i) {
      const int b = i + 1;
      if (!joints.count(joint_names[i])) throw std::runtime_error("URDF has no joint " + joint_names[i]);
      const Xml* j = joints[joint_names[i]];
      const std::string child = j->child("child")->get("link"), parent = j->child("parent")->get("link");
      if (child != body_names[b]) throw std::runtime_error("URDF joint " + joint_names[i] + " drives " + child + ", expected " + body_names[b]);
      int pb = -1;
      for (int k = 0; k < b; ++k) if (body_names[k] == parent) pb = k;
      if (pb < 0) throw std::runtime_error("parent link of " + joint_names[i] + " does not precede it");
      m.parent[b] = pb;
      const Xml* o = j->child("origin");
      auto xyz = numbers(o ? o->get("xyz", "0 0 0") : "0 0 0", 3, "joint xyz"), rpy = numbers(o ? o->get("rpy", "0 0 0") : "0 0 0", 3, "joint rpy");
      for (int k = 0; k < 3; ++k) m.pos[b][k] = xyz[k];
      rpy2mat(rpy, m.rfix[b]);
      m.has_rfix[b] = (rpy[0] != 0.0 || rpy[1] != 0.0 || rpy[2] != 0.0) ? 1 : 0;
      m.axis[b] = axis_index(numbers(j->child("axis")->get("xyz"), 3, "joint axis"));
      const Xml* lim = j->child("limit");
      if (!lim) throw std::runtime_error("joint without <limit>: " + joint_names[i]);
      m.jnt_range[i][0] = std::stod(lim->get("lower")); m.jnt_range[i][1] = std::stod(lim->get("upper"));
      m.ctrl_range[i][0] = -std::stod(lim->get("effort")); m.ctrl_range[i][1] = std::stod(lim->get("effort"));
      inertial(child, b);
    }
    m.foot_body[0] = m.foot_body[1] = -1;
    for (int b = 0; b < H1_NB; ++b) {
      if (body_names[b] == "left_ankle_link") m.foot_body[0] = b;
      if (body_names[b] == "right_ankle_link") m.foot_body[1] = b;
    }
    m.total_mass = 0.0;
    for (int b = 0; b < H1_NB; ++b) m.total_mass += m.mass[b];
    *out = m;
    return true;
  } catch (const std::exception& e) {
    if (error) *error = e.what();
    return false;
  }
}
