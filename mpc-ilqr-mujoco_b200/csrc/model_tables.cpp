// Host-side construction of the device model tables (h1::DynModel / h1::CostModel) from the plain-C
// H1Model description: ancestor lists, dof-tree levels and subtree ranges used by the lane-parallel kernels.
#include <cstdlib>
#include "model_tables.h"
#include "../../include/h1_model_data.h"
#include <cstring>

extern "C" const H1Model* h1_default_dynamics_model(void) { return &H1_DYNAMICS_MODEL; }
extern "C" const H1Model* h1_default_cost_model(void) { return &H1_COST_MODEL; }

namespace h1 {

static void tree_common(const H1Model& m, int* parent, int* axis, int* has_rfix, int* depth, int (*anc_body)[6],
                        int* chain_end) {
  for (int b = 0; b < NB; ++b) {
    parent[b] = m.parent[b]; axis[b] = m.axis[b] < 0 ? 0 : m.axis[b]; has_rfix[b] = m.has_rfix[b];
    depth[b] = (b == 0) ? 0 : depth[m.parent[b]] + 1;
  }
  for (int b = 0; b < NB; ++b) {
    for (int d = 0; d < 6; ++d) anc_body[b][d] = 0;
    for (int a = b; a >= 0; a = m.parent[a]) anc_body[b][depth[a]] = a;
    chain_end[b] = b;
  }
  for (int b = NB - 1; b >= 1; --b)
    if (chain_end[b] > chain_end[m.parent[b]]) chain_end[m.parent[b]] = chain_end[b];
}

bool build_dyn_model(const H1Model& m, DynModel* d) {
  std::memset(d, 0, sizeof(*d));
  for (int b = 0; b < NB; ++b) {
    if (b > 0 && (m.parent[b] < 0 || m.parent[b] >= b)) return false;  // DFS order required
    for (int i = 0; i < 3; ++i) { d->pos[b][i] = m.pos[b][i]; d->ipos[b][i] = m.ipos[b][i]; }
    for (int i = 0; i < 9; ++i) d->rfix[b][i] = m.rfix[b][i];
    for (int i = 0; i < 6; ++i) d->inertia[b][i] = m.inertia[b][i];
    d->mass[b] = m.mass[b];
  }
  tree_common(m, d->parent, d->axis, d->has_rfix, d->depth, d->anc_body, d->chain_end);
  for (int b = 0; b < NB; ++b) if (d->depth[b] > 5) return false;
  for (int j = 0; j < NV; ++j) { d->armature[j] = m.armature[j]; d->damping[j] = m.damping[j]; }
  for (int i = 0; i < NU; ++i) {
    d->ctrl_lo[i] = m.ctrl_range[i][0]; d->ctrl_hi[i] = m.ctrl_range[i][1];
    d->jnt_lo[i] = m.jnt_range[i][0]; d->jnt_hi[i] = m.jnt_range[i][1];
  }
  for (int f = 0; f < H1_NFOOT; ++f) {
    d->foot_body[f] = m.foot_body[f];
    d->foot_dof[f] = 5 + m.foot_body[f];
    for (int c = 0; c < H1_NCP; ++c)
      for (int i = 0; i < 3; ++i) d->foot_pts[f * H1_NCP + c][i] = m.foot_pts[f][c][i];
  }
  for (int i = 0; i < 3; ++i) d->gravity[i] = m.gravity[i];
  d->h = m.timestep; d->kn = m.contact_kn; d->bn = m.contact_bn; d->bt = m.contact_bt; d->eps = m.contact_eps;
  d->total_mass = m.total_mass;
  d->inv_total_mass = 1.0 / m.total_mass;
  // dof tree: base dofs 0..5 form a chain, hinge dof 5+b hangs below its parent body's last dof
  for (int j = 0; j < NV; ++j) {
    int n = 0;
    if (j < 6) { for (int k = 0; k <= j; ++k) d->alist[j][n++] = k; }
    else {
      int b = j - 5;
      for (int k = 0; k < 6; ++k) d->alist[j][n++] = k;
      for (int dd = 1; dd <= d->depth[b]; ++dd) d->alist[j][n++] = 5 + d->anc_body[b][dd];
    }
    if (n > MAXSLOT) return false;
    d->nlist[j] = n;
    d->level[j] = n - 1;
    d->dof_sub_end[j] = (j < 6) ? NV - 1 : 5 + d->chain_end[j - 5];
  }
  // each foot must hang on a full-depth chain so that its dof list has MAXSLOT entries
  for (int f = 0; f < H1_NFOOT; ++f) if (d->nlist[d->foot_dof[f]] != MAXSLOT) return false;
  for (int j = 0; j < NV; ++j) { d->cp_lo[j] = NCPT; d->cp_hi[j] = 0; }
  for (int f = 0; f < H1_NFOOT; ++f)
    for (int s = 0; s < MAXSLOT; ++s) {
      int j = d->alist[d->foot_dof[f]][s];
      if (f * H1_NCP < d->cp_lo[j]) d->cp_lo[j] = f * H1_NCP;
      if ((f + 1) * H1_NCP > d->cp_hi[j]) d->cp_hi[j] = (f + 1) * H1_NCP;
    }
  for (int j = 0; j < NV; ++j) if (d->cp_hi[j] == 0) d->cp_lo[j] = 0;
  d->n_base_children = 0;
  for (int b = 1; b < NB; ++b)
    if (m.parent[b] == 0) { d->base_child_slot[b] = d->n_base_children++; }
  if (d->n_base_children > 4) return false;
  for (int b = 1; b < NB; ++b) d->nchild[m.parent[b]]++;
  for (int b = 1; b < NB; ++b)
    if (d->nchild[b] > 1 && d->depth[b] >= 3) return false;  // SEQ_MAXSAVE (h1_lin_dirs.cuh)
  {  // hinged bodies by decreasing subtree size (stable): work order of the direction-uniform linearization
    int n = 0;
    for (int sz = NB; sz >= 1; --sz)
      for (int b = 1; b < NB; ++b)
        if (d->chain_end[b] - b + 1 == sz) d->dir_order[n++] = b;
  }
  {  // merged tangent kernel (h1_lin_finish.cuh): the 49 direction items of a knot group, class by class — 2 (base rotations, z and
     // base linear velocity, then the quaternion-Jacobian item), 0 (hinge angles), 1 (base angular velocity, hinge rates) — and
     // costliest first within a class (hinges in dir_order). Measured on B200 per 8192-instance linearization (tools/lint_prof.py):
     // one launch per class 8.83 ms (the eight warps of a CTA are busy 82 % of the time), this order 8.26 ms (92 %), 0-2-1 8.33,
     // 1-0-2 8.55, all items sorted by measured cost 8.49: an item runs ~20 % slower when all eight warps are busy (the 2 KB
     // stack frames of 256 threads no longer fit L1), and mixing the classes' code in time costs more than the last 3 % of balance.
    const int nh = NB - 1;
    const char* seq = "201";
    if (const char* e = getenv("H1_TAN_ORDER")) seq = e;   // A/B measurements: another class sequence
    int k = 0;
    for (const char* c = seq; *c && k <= 49; ++c) {
      const int cls = *c - '0';
      if (cls == 0) for (int i = 0; i < nh; ++i) d->tan_order[k++] = (unsigned char)((0 << 5) | i);
      if (cls == 1) for (int i = 0; i < 3 + nh; ++i) d->tan_order[k++] = (unsigned char)((1 << 5) | i);
      if (cls == 2) { const int o[8] = {0, 1, 2, 4, 5, 6, 7, 3}; for (int i = 0; i < 8; ++i) d->tan_order[k++] = (unsigned char)((2 << 5) | o[i]); }
    }
    if (k != 2 * nh + 11) return false;
    d->n_tan_items = k;
  }
  // thread-sequential f_D (h1_dyn_seq.cuh): specialised for base + two 5-hinge leg chains ending in the feet
  // (bodies 1-5, 6-10) + torso (11) + two 4-hinge arm chains below the torso (12-15, 16-19)
  {
    bool ok = d->foot_body[0] == 5 && d->foot_body[1] == 10;
    for (int b = 1; b < NB && ok; ++b) {
      int want = b - 1;
      if (b == 1 || b == 6 || b == 11) want = 0;
      if (b == 12 || b == 16) want = 11;
      ok = m.parent[b] == want;
    }
    for (int k = 0; k < NV && ok; ++k) {   // the compile-time dof tree (h1_common.cuh) must be this model's
      ok = d->nlist[k] == h1_nlist(k);
      for (int sl = 0; sl < d->nlist[k] && ok; ++sl) ok = d->alist[k][sl] == h1_anc(k, sl);
    }
    d->seq_ok = ok ? 1 : 0;
  }
  return true;
}

bool build_cost_model(const H1Model& m, CostModel* c) {
  std::memset(c, 0, sizeof(*c));
  for (int b = 0; b < NB; ++b) {
    if (b > 0 && (m.parent[b] < 0 || m.parent[b] >= b)) return false;
    for (int i = 0; i < 3; ++i) { c->pos[b][i] = m.pos[b][i]; c->ipos[b][i] = m.ipos[b][i]; }
    for (int i = 0; i < 9; ++i) c->rfix[b][i] = m.rfix[b][i];
    c->wmass[b] = m.mass[b] / m.total_mass;
  }
  tree_common(m, c->parent, c->axis, c->has_rfix, c->depth, c->anc_body, c->chain_end);
  for (int f = 0; f < H1_NFOOT; ++f) c->foot_body[f] = m.foot_body[f];
  // joint pairs k <= l: ancestor-or-self pairs first
  int n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = 1; k < NB; ++k)
      for (int l = k; l < NB; ++l) {
        const bool anc = c->depth[k] <= c->depth[l] && c->anc_body[l][c->depth[k]] == k;
        if (anc == (pass == 0)) { c->pair_k[n] = (unsigned char)k; c->pair_l[n] = (unsigned char)l; ++n; }
      }
    if (pass == 0) c->n_anc_pairs = n;
  }
  return true;
}

}  // namespace h1
