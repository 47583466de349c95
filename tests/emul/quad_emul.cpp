// Emulation build of the quad-cooperative f_D (mpc-ilqr-mujoco_b200/csrc/h1_dyn_quad.cuh) for the CPU test suite:
// the four lanes of an evaluation run as four host threads, the xor-shuffles of the kernel become exchanges through
// a small shared array between two barriers. Same source as the kernel; test infrastructure only.
#include "../../mpc-ilqr-mujoco_b200/csrc/h1_dyn_quad.cuh"
#include "../../mpc-ilqr-mujoco_b200/csrc/model_tables.h"
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

namespace {
struct Barrier {
  std::mutex m; std::condition_variable cv; int count = 0, gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int g = gen;
    if (++count == 4) { count = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct QuadHost {
  Barrier* bar; double* slot; int g;
  double exch(double v, int mask) const {
    slot[g] = v;
    bar->wait();
    const double o = slot[g ^ mask];
    bar->wait();
    return o;
  }
  double xor1(double v) const { return exch(v, 1); }
  double xor2(double v) const { return exch(v, 2); }
  void sync() const { bar->wait(); }
};
}  // namespace

// x_next = f_D(x, u) and the dynamics-model CoM for n states through dyn_step_quad; com2 = dyn_com_quad of the same states
extern "C" int emul_dyn_step_quad(int n, const double* x, const double* u, double* xn, double* com, double* com2) {
  static h1::DynModel md;
  static bool init = false;
  if (!init) { if (!h1::build_dyn_model(*h1_default_dynamics_model(), &md)) return -1; init = true; }
  if (!md.seq_ok) return -2;
  Barrier bar;
  double slot[4];
  double store[h1::Q4_STORE * 4];
  auto lane = [&](int g) {
    QuadHost cx{&bar, slot, g};
    for (int i = 0; i < n; ++i) {
      const double* xi = x + (size_t)i * h1::NX;
      double qn[h1::Q4_CHAIN], vn[h1::Q4_CHAIN], bn[13], c[3], c2[3];
      h1::dyn_step_quad(md, cx, g, xi, u ? u + (size_t)i * h1::NU : nullptr, store + g, 4, qn, vn, bn, c);
      h1::dyn_com_quad(md, cx, g, xi, c2);
      double* o = xn + (size_t)i * h1::NX;
      for (int k = 0; k < h1::Q4_CHAIN; ++k) {
        if (g == 3 && k == 0) continue;   // the torso is written by lane 2
        const int b = h1::q4_body(g, k);
        o[6 + b] = qn[k]; o[h1::NQ + 5 + b] = vn[k];
      }
      if (g == 0) {
        for (int k = 0; k < 7; ++k) o[k] = bn[k];
        for (int k = 0; k < 6; ++k) o[h1::NQ + k] = bn[7 + k];
        for (int k = 0; k < 3; ++k) { com[3 * i + k] = c[k]; com2[3 * i + k] = c2[k]; }
      }
      cx.sync();
    }
  };
  std::thread t1(lane, 1), t2(lane, 2), t3(lane, 3);
  lane(0);
  t1.join(); t2.join(); t3.join();
  return 0;
}
