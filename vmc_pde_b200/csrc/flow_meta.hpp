// Host-side construction of FlowMeta (offsets of the flat parameter vector in the reference's order,
// var_state.py:106-108 over flax's string-sorted param dict) from the public config.
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/vmcpde.h"
#include "flow_core.cuh"

namespace vmc {

inline int make_flow_meta(const vmcpde_flow_config* c, FlowMeta* m, std::string* err) {
  auto fail = [&](int code, const char* msg) { if (err) *err = msg; return code; };
  if (!c || !m) return fail(VMCPDE_EINVAL, "null config");
  if (c->dim < 2 || c->dim > kMaxDim) return fail(VMCPDE_EUNSUPPORTED, "dim must be in [2,16]");
  if (c->depth < 0 || c->depth > kMaxDepth) return fail(VMCPDE_EUNSUPPORTED, "depth must be in [0,32]");
  if (c->depth > 0 && (c->n_hidden_layers < 1 || c->n_hidden_layers > kMaxLayers))
    return fail(VMCPDE_EUNSUPPORTED, "a SingleTrafo has 1 to 3 hidden layers in this build");
  if (c->depth > 0 && c->n_hidden_layers > 1 && !c->hidden_widths)
    return fail(VMCPDE_EINVAL, "hidden_widths is required when n_hidden_layers > 1");
  int widths[kMaxLayers] = {c->hidden, 0, 0};
  if (c->depth > 0 && c->hidden_widths)
    for (int l = 0; l < c->n_hidden_layers; ++l) widths[l] = c->hidden_widths[l];
  if (c->depth > 0 && c->n_hidden_layers == 1 && (widths[0] < 1 || widths[0] > kMaxHidden))
    return fail(VMCPDE_EUNSUPPORTED, "hidden width must be in [1,256]");
  if (c->depth > 0 && c->n_hidden_layers > 1) {   // generic path: per-layer activations (and their jets) live in thread-local arrays
    int sum = 0;
    for (int l = 0; l < c->n_hidden_layers; ++l) {
      if (widths[l] < 1 || widths[l] > kMLWidth) return fail(VMCPDE_EUNSUPPORTED, "with several hidden layers every width must be in [1,32]");
      sum += widths[l];
    }
    if (sum > kMaxHidden) return fail(VMCPDE_EUNSUPPORTED, "sum of hidden widths must be <= 256");
  }
  const int variant = c->variant & 0xff, gc = (c->variant & VMCPDE_GLOBAL_CHANGE) ? 1 : 0;
  if (c->variant < 0 || variant > 3 || (c->variant & ~(0xff | VMCPDE_GLOBAL_CHANGE))) return fail(VMCPDE_EINVAL, "bad coupling variant");
  if (c->latent < 0 || c->latent > 1) return fail(VMCPDE_EINVAL, "bad latent distribution");
  if (c->depth > 0 && c->dim < 2) return fail(VMCPDE_EINVAL, "coupling blocks need dim >= 2");
  const int d = c->dim, d1 = d / 2, d2 = d - d / 2;
  *m = FlowMeta{};
  m->d = d; m->depth = c->depth; m->h = c->depth > 0 ? widths[0] : 1; m->variant = variant; m->latent = c->latent;
  m->gc = c->depth > 0 ? gc : 0;
  m->nl = c->depth > 0 ? c->n_hidden_layers : 1;
  for (int l = 0; l < kMaxLayers; ++l) m->hw[l] = (c->depth > 0 && l < m->nl) ? widths[l] : (l == 0 ? 1 : 0);
  int off = 0;
  m->off_L = off; off += d * (d - 1) / 2;
  m->off_Ldiag = off; off += d;
  m->off_dist = off; off += (c->latent == VMCPDE_STUDENT_T) ? 1 : 0;
  m->off_mu = off; off += d;
  const int T1 = trafo_size(d1, d2, *m), T2 = trafo_size(d2, d1, *m);
  const int per_block = (variant == VMCPDE_DIFFERENT_ADD ? 2 : 1) * (T1 + T2) + (m->gc ? d + 1 : 0);
  std::vector<int> order(c->depth);
  for (int b = 0; b < c->depth; ++b) order[b] = b;
  std::sort(order.begin(), order.end(), [](int a, int b) {
    return ("blocks_" + std::to_string(a)) < ("blocks_" + std::to_string(b));
  });
  for (int b : order) { m->block_off[b] = off; off += per_block; }
  m->P = off;
  for (int b = 0; b < c->depth; ++b) {
    std::vector<int> seen(d, 0);
    for (int i = 0; i < d1; ++i) {
      int v = c->ind_up[b * d1 + i];
      if (v < 0 || v >= d || seen[v]++) return fail(VMCPDE_EINVAL, "ind_up/ind_down must partition range(dim)");
      m->up[b][i] = (int8_t)v;
    }
    for (int i = 0; i < d2; ++i) {
      int v = c->ind_down[b * d2 + i];
      if (v < 0 || v >= d || seen[v]++) return fail(VMCPDE_EINVAL, "ind_up/ind_down must partition range(dim)");
      m->down[b][i] = (int8_t)v;
    }
  }
  for (int i = 0; i < d; ++i) m->offset[i] = c->offset ? c->offset[i] : 0.0;
  return VMCPDE_OK;
}

// dispatch a callable templated on the compile-time dimension
#define VMC_DISPATCH_DIM(dim, ...)                                   \
  switch (dim) {                                                     \
    case 2: { constexpr int D = 2; __VA_ARGS__; break; }             \
    case 3: { constexpr int D = 3; __VA_ARGS__; break; }             \
    case 4: { constexpr int D = 4; __VA_ARGS__; break; }             \
    case 5: { constexpr int D = 5; __VA_ARGS__; break; }             \
    case 6: { constexpr int D = 6; __VA_ARGS__; break; }             \
    case 8: { constexpr int D = 8; __VA_ARGS__; break; }             \
    case 10: { constexpr int D = 10; __VA_ARGS__; break; }           \
    case 12: { constexpr int D = 12; __VA_ARGS__; break; }           \
    default: return VMCPDE_EUNSUPPORTED;                             \
  }

}  // namespace vmc
