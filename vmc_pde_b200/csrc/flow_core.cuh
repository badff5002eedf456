// Per-sample math of the vmc_pde ansatz: RealNVP-style flow + latent log-pdf, its forward jets
// (value, directional first derivatives, weighted sum of directional second derivatives) and a
// memory-free reverse sweep that emits O = d logp / d theta in the reference's flat order.
//
// Reference semantics restated here (not copied): net.py:10-36 (latent pdfs), :44-61 (SingleTrafo),
// :84-153 (SingleBlock fwd/inv), :168-182 (INN), :209-217 (INNwProb); util.py:21-26 (covariance);
// var_state.py:31-32 (value_and_grad / jacrev(jacfwd) replaced by hand-derived jets);
// evolutionEq.py:84-119 (local terms).  Flat layout: var_state.py:106-108 over flax's sorted dict.
//
// Everything is __host__ __device__ so tests/hostsim can instantiate the same code on the CPU
// (test tooling only; libvmcpde.so exports no CPU entry point for it).
#pragma once
#include <cstdint>
#include <cmath>

#ifdef __CUDACC__
#define VMC_HD __host__ __device__ __forceinline__
#define VMC_HD_COLD __host__ __device__ __noinline__   // rarely taken paths kept out of line: the one-layer fast path keeps its registers
#else
#define VMC_HD inline
#define VMC_HD_COLD inline
#endif

namespace vmc {

constexpr int kMaxDepth = 32;
constexpr int kMaxDim = 16;
constexpr int kMaxHalf = 8;
constexpr int kMaxHidden = 256;
constexpr double kAlpha = 10.0;  // net.py:50

enum Variant { kNoAdd = 0, kDifferentAdd = 1, kJacEq1 = 2, kAddS = 3 };  // net.py:69-71 + else branch
enum Latent { kGauss = 0, kStudentT = 1 };                                // net.py:197-198
enum Equation {                                                             // evolutionEq.py:54-60
  kDiffusion = 0, kDiffusionDrift = 1, kDiffusionAniso = 2,
  kAdvectionHamiltonian = 3, kAdvectionPaper = 4, kAdvectionHamiltonianWDiss = 5
};

// VMC_ML: the per-dimension kernels are compiled twice, without (0) and with (1) the generic multi-layer SingleTrafo path, so the
// one-hidden-layer kernels keep the registers and stack frame they had before the generic path existed.
#ifndef VMC_ML
#define VMC_ML 1
#endif
constexpr int kMaxLayers = 3;    // hidden layers per SingleTrafo (net.py:53-58 loops over `intmediate`)
constexpr int kMLWidth = 32;     // widest hidden layer of the multi-layer path (its jets live in per-thread arrays)

struct FlowMeta {
  int d, depth, h, variant, latent, P;
  int nl, hw[kMaxLayers];          // hidden layers and their widths (hw[0] == h); nl == 1 is the streaming fast path
  int gc;                          // SingleBlock.global_change (net.py:72,80-82): per block global_offset[d], global_scale[1]
  int off_L, off_Ldiag, off_dist, off_mu;
  int block_off[kMaxDepth];        // start of "blocks_b" in the flat vector (string-sorted order)
  int8_t up[kMaxDepth][kMaxHalf];   // ind_up   (d/2 entries)
  int8_t down[kMaxDepth][kMaxHalf]; // ind_down (d - d/2 entries)
  double offset[kMaxDim];           // network_args["offset"]
};

struct EqParams {  // evolutionEq.py:61-77
  int mode;
  double D, mu, m, omega, lam, T, gamma, t;
};

// ------------------------------------------------------------------------------------------------
struct Trafo {  // one SingleTrafo: Dense_0{bias,kernel} ... Dense_nl{bias,kernel} (net.py:44-61)
  const double *b1, *W1, *b2, *W2;   // the two layers of the single-hidden-layer fast path
  int h;
  int nl, hw[kMaxLayers];            // nl > 1: generic path, layers walked from p0
  const double* p0;
};
VMC_HD int trafo_size(int din, int dout, int h) { return h + din * h + dout + h * dout; }
VMC_HD Trafo trafo_at(const double* th, int off, int din, int dout, int h) {
  Trafo t;
  t.b1 = th + off; t.W1 = t.b1 + h; t.b2 = t.W1 + din * h; t.W2 = t.b2 + dout; t.h = h;
  t.nl = 1; t.hw[0] = h; t.p0 = th + off;
  return t;
}
VMC_HD int trafo_size(int din, int dout, const FlowMeta& m) {
  if (!VMC_ML || m.nl == 1) return trafo_size(din, dout, m.h);
  int s = 0, in = din;
  for (int l = 0; l < m.nl; ++l) { s += m.hw[l] + in * m.hw[l]; in = m.hw[l]; }
  return s + dout + in * dout;
}
VMC_HD Trafo trafo_at(const double* th, int off, int din, int dout, const FlowMeta& m) {
  Trafo t = trafo_at(th, off, din, dout, m.h);
#if VMC_ML
  t.nl = m.nl;
  for (int l = 0; l < kMaxLayers; ++l) t.hw[l] = l < m.nl ? m.hw[l] : 0;
#endif
  return t;
}

// Blocks with global_change start with global_offset[d], global_scale[1] ("global_*" sorts before "s1"); the trafos follow.
VMC_HD bool has_global_change(const FlowMeta& m) { return VMC_ML && m.gc; }
VMC_HD int block_base(const FlowMeta& m, int b) { return m.block_off[b] + (has_global_change(m) ? m.d + 1 : 0); }

// ---- generic multi-layer path (nl > 1): layer l maps prev (nprev) -> hw[l] with tanh, the last layer -> NO with alpha*tanh.
// Parameters of layer l: bias[hw[l]], kernel[nprev][hw[l]] (row-major, flax Dense).  `hid` receives the activations of all
// hidden layers, one after the other (sum of widths <= kMaxHidden).
template <int NI, int NO>
VMC_HD_COLD void trafo_value_ml(const Trafo& t, const double* a0, double* out, double* hid) {
  double buf[2][kMLWidth];
  const double* prev = a0;
  const double* p = t.p0;
  int nprev = NI, hoff = 0;
  for (int l = 0; l < t.nl; ++l) {
    const int h = t.hw[l];
    const double *b = p, *W = p + h;
    double* cur = buf[l & 1];
    for (int j = 0; j < h; ++j) {
      double s = b[j];
      for (int i = 0; i < nprev; ++i) s = fma(W[i * h + j], prev[i], s);
      cur[j] = tanh(s);
      if (hid) hid[hoff + j] = cur[j];
    }
    p += h + nprev * h; prev = cur; nprev = h; hoff += h;
  }
  const double *b = p, *W = p + NO;
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    double s = b[o];
    for (int j = 0; j < nprev; ++j) s = fma(W[j * NO + o], prev[j], s);
    out[o] = kAlpha * tanh(s);
  }
}

// value-only forward; optionally keeps the hidden activations for the reverse sweep
template <int NI, int NO>
VMC_HD void trafo_value(const Trafo& t, const double* a0, double* out, double* hid) {
  if (VMC_ML && t.nl > 1) { trafo_value_ml<NI, NO>(t, a0, out, hid); return; }
  double acc[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) acc[o] = t.b2[o];
  for (int j = 0; j < t.h; ++j) {
    double p = t.b1[j];
#pragma unroll
    for (int i = 0; i < NI; ++i) p = fma(t.W1[i * t.h + j], a0[i], p);
    const double a = tanh(p);
    if (hid) hid[j] = a;
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[o] = fma(t.W2[j * NO + o], a, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) out[o] = kAlpha * tanh(acc[o]);
}

// ------------------------------------------------------------------------------------------------
// Jets: value, D directional first derivatives, and l = sum_k w_k d^2/dv_k^2.
template <int D>
struct Jet {
  double v, g[D], l;
};
template <int D> VMC_HD void jet_const(Jet<D>& a, double v) {
  a.v = v; a.l = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) a.g[k] = 0.0;
}
template <int D> VMC_HD void jet_axpy(Jet<D>& y, double c, const Jet<D>& x) {
  y.v = fma(c, x.v, y.v); y.l = fma(c, x.l, y.l);
#pragma unroll
  for (int k = 0; k < D; ++k) y.g[k] = fma(c, x.g[k], y.g[k]);
}
template <int D> VMC_HD void jet_add(Jet<D>& y, const Jet<D>& x) { jet_axpy(y, 1.0, x); }
// out = scale * tanh(p)
template <int D> VMC_HD void jet_tanh(const Jet<D>& p, const double* w, double scale, Jet<D>& out) {
  const double y = tanh(p.v), y1 = 1.0 - y * y, y2 = -2.0 * y * y1;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) s = fma(w[k] * p.g[k], p.g[k], s);
  out.l = scale * (y1 * p.l + y2 * s);
#pragma unroll
  for (int k = 0; k < D; ++k) out.g[k] = scale * y1 * p.g[k];
  out.v = scale * y;
}
// u <- u * exp(s)
template <int D> VMC_HD void jet_mulexp(Jet<D>& u, const Jet<D>& s, const double* w) {
  const double e = exp(s.v);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) acc += w[k] * (2.0 * u.g[k] * s.g[k] + u.v * s.g[k] * s.g[k]);
  u.l = e * (u.l + u.v * s.l + acc);
#pragma unroll
  for (int k = 0; k < D; ++k) u.g[k] = e * (u.g[k] + u.v * s.g[k]);
  u.v = u.v * e;
}

template <int D, int NI, int NO>
VMC_HD_COLD void trafo_jet_ml(const Trafo& t, const Jet<D>* a0, const double* w, Jet<D>* out) {
  Jet<D> buf[2][kMLWidth];
  const Jet<D>* prev = a0;
  const double* p = t.p0;
  int nprev = NI;
  for (int l = 0; l < t.nl; ++l) {
    const int h = t.hw[l];
    const double *b = p, *W = p + h;
    Jet<D>* cur = buf[l & 1];
    for (int j = 0; j < h; ++j) {
      Jet<D> q;
      jet_const(q, b[j]);
      for (int i = 0; i < nprev; ++i) jet_axpy(q, W[i * h + j], prev[i]);
      jet_tanh(q, w, 1.0, cur[j]);
    }
    p += h + nprev * h; prev = cur; nprev = h;
  }
  const double *b = p, *W = p + NO;
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    Jet<D> q;
    jet_const(q, b[o]);
    for (int j = 0; j < nprev; ++j) jet_axpy(q, W[j * NO + o], prev[j]);
    jet_tanh(q, w, kAlpha, out[o]);
  }
}

template <int D, int NI, int NO>
VMC_HD void trafo_jet(const Trafo& t, const Jet<D>* a0, const double* w, Jet<D>* out) {
  if (VMC_ML && t.nl > 1) { trafo_jet_ml<D, NI, NO>(t, a0, w, out); return; }
  Jet<D> acc[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) jet_const(acc[o], t.b2[o]);
  for (int j = 0; j < t.h; ++j) {
    Jet<D> p, a;
    jet_const(p, t.b1[j]);
#pragma unroll
    for (int i = 0; i < NI; ++i) jet_axpy(p, t.W1[i * t.h + j], a0[i]);
    jet_tanh(p, w, 1.0, a);
#pragma unroll
    for (int o = 0; o < NO; ++o) jet_axpy(acc[o], t.W2[j * NO + o], a);
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) jet_tanh(acc[o], w, kAlpha, out[o]);
}

// ------------------------------------------------------------------------------------------------
// Reverse sweep through one trafo.  dout = d logp / d(trafo output).  Emits the parameter gradients in
// flat order (Dense_0/bias, Dense_0/kernel, Dense_1/bias, Dense_1/kernel) when EMIT, and accumulates
// d logp / d(input) into din_acc when INGRAD.  `hid` holds the hidden activations from trafo_value.
// generic multi-layer reverse: the deltas of all hidden layers go to dh (same offsets as hid), then the gradients are
// emitted layer by layer in flat order
template <int NI, int NO, bool EMIT, bool INGRAD, class Emit>
VMC_HD_COLD void trafo_reverse_ml(const Trafo& t, const double* a0, const double* out, const double* dout,
                             const double* hid, double* dh, double* din_acc, Emit& em) {
  double dp2[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    const double y = out[o] * (1.0 / kAlpha);
    dp2[o] = dout[o] * kAlpha * (1.0 - y * y);
  }
  // parameter / activation offsets of the layers
  int poff[kMaxLayers + 1], hoff[kMaxLayers + 1], nin[kMaxLayers + 1];
  {
    int po = 0, ho = 0, in = NI;
    for (int l = 0; l < t.nl; ++l) { poff[l] = po; hoff[l] = ho; nin[l] = in; po += t.hw[l] + in * t.hw[l]; ho += t.hw[l]; in = t.hw[l]; }
    poff[t.nl] = po; hoff[t.nl] = ho; nin[t.nl] = in;
  }
  {  // last hidden layer from the output layer
    const int L = t.nl - 1, h = t.hw[L];
    const double* W = t.p0 + poff[t.nl] + NO;
    for (int j = 0; j < h; ++j) {
      double s = 0.0;
#pragma unroll
      for (int o = 0; o < NO; ++o) s = fma(W[j * NO + o], dp2[o], s);
      const double a = hid[hoff[L] + j];
      dh[hoff[L] + j] = (1.0 - a * a) * s;
    }
  }
  for (int l = t.nl - 2; l >= 0; --l) {  // earlier hidden layers
    const int h = t.hw[l], hn = t.hw[l + 1];
    const double* W = t.p0 + poff[l + 1] + hn;     // kernel of layer l + 1: [h][hn]
    for (int i = 0; i < h; ++i) {
      double s = 0.0;
      for (int j = 0; j < hn; ++j) s = fma(W[i * hn + j], dh[hoff[l + 1] + j], s);
      const double a = hid[hoff[l] + i];
      dh[hoff[l] + i] = (1.0 - a * a) * s;
    }
  }
  if (EMIT) {
    for (int l = 0; l < t.nl; ++l) {
      const int h = t.hw[l];
      for (int j = 0; j < h; ++j) em.put(dh[hoff[l] + j]);                         // Dense_l/bias
      for (int i = 0; i < nin[l]; ++i) {                                           // Dense_l/kernel [nin][h]
        const double ai = l == 0 ? a0[i] : hid[hoff[l - 1] + i];
        for (int j = 0; j < h; ++j) em.put(ai * dh[hoff[l] + j]);
      }
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) em.put(dp2[o]);                                   // Dense_nl/bias
    const int L = t.nl - 1;
    for (int j = 0; j < t.hw[L]; ++j) {
#pragma unroll
      for (int o = 0; o < NO; ++o) em.put(hid[hoff[L] + j] * dp2[o]);              // Dense_nl/kernel [h][NO]
    }
  }
  if (INGRAD) {
    const int h = t.hw[0];
    const double* W = t.p0 + h;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      double s = 0.0;
      for (int j = 0; j < h; ++j) s = fma(W[i * h + j], dh[j], s);
      din_acc[i] += s;
    }
  }
}

template <int NI, int NO, bool EMIT, bool INGRAD, class Emit>
VMC_HD void trafo_reverse(const Trafo& t, const double* a0, const double* out, const double* dout,
                          const double* hid, double* dh, double* din_acc, Emit& em) {
  if (VMC_ML && t.nl > 1) { trafo_reverse_ml<NI, NO, EMIT, INGRAD>(t, a0, out, dout, hid, dh, din_acc, em); return; }
  double dp2[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    const double y = out[o] * (1.0 / kAlpha);
    dp2[o] = dout[o] * kAlpha * (1.0 - y * y);
  }
  for (int j = 0; j < t.h; ++j) {
    double s = 0.0;
#pragma unroll
    for (int o = 0; o < NO; ++o) s = fma(t.W2[j * NO + o], dp2[o], s);
    const double v = (1.0 - hid[j] * hid[j]) * s;
    dh[j] = v;
    if (EMIT) em.put(v);
  }
  if (EMIT) {
#pragma unroll
    for (int i = 0; i < NI; ++i)
      for (int j = 0; j < t.h; ++j) em.put(a0[i] * dh[j]);
#pragma unroll
    for (int o = 0; o < NO; ++o) em.put(dp2[o]);
    for (int j = 0; j < t.h; ++j) {
#pragma unroll
      for (int o = 0; o < NO; ++o) em.put(hid[j] * dp2[o]);
    }
  }
  if (INGRAD) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      double s = 0.0;
      for (int j = 0; j < t.h; ++j) s = fma(t.W1[i * t.h + j], dh[j], s);
      din_acc[i] += s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
VMC_HD double digamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                   f * (-5.0 / 660.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

VMC_HD int triu_index(int d, int a, int b) { return a * d - (a * (a + 1)) / 2 + (b - a - 1); }

// L (upper triangular, util.py:21-26) from theta
template <int D>
VMC_HD void build_L(const FlowMeta& m, const double* th, double (&L)[D][D]) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
#pragma unroll
    for (int b = 0; b < D; ++b) L[a][b] = 0.0;
    L[a][a] = exp(th[m.off_Ldiag + a]);
#pragma unroll
    for (int b = a + 1; b < D; ++b) L[a][b] = th[m.off_L + triu_index(D, a, b)];
  }
}

// latent log-pdf of y = z - offset (mu subtracted here), net.py:14-20 / 27-36 via |L^-1 y|^2
template <int D>
VMC_HD double latent_logpdf(const FlowMeta& m, const double* th, const double* y_in) {
  double L[D][D];
  build_L<D>(m, th, L);
  double w[D];
  double q = 0.0, sumLd = 0.0;
#pragma unroll
  for (int a = D - 1; a >= 0; --a) {
    double s = y_in[a] - th[m.off_mu + a];
#pragma unroll
    for (int b = a + 1; b < D; ++b) s -= L[a][b] * w[b];
    w[a] = s / L[a][a];
    q += w[a] * w[a];
    sumLd += th[m.off_Ldiag + a];
  }
  if (m.latent == kGauss) return -0.5 * (D * 1.8378770664093454835606594728112 + 2.0 * sumLd + q);
  const double nu = exp(th[m.off_dist]) + 1.0;
  return lgamma(0.5 * (nu + D)) - lgamma(0.5 * nu) - 0.5 * D * log(nu * 3.14159265358979323846) -
         0.5 * (nu + D) * log(1.0 + q / nu);
}

// ------------------------------------------------------------------------------------------------
// value-only block maps (net.py:84-153) on the full coordinate vector z; returns the block's log-Jacobian
template <int D>
VMC_HD double block_forward_value(const FlowMeta& m, const double* th, int b, double* z) {
  constexpr int D1 = D / 2, D2 = D - D / 2;
  const int T1 = trafo_size(D1, D2, m), T2 = trafo_size(D2, D1, m);
  const int o = block_base(m, b);
  double u1[D1], u2[D2], s2[D1], s1[D2], lj = 0.0;
#pragma unroll
  for (int i = 0; i < D1; ++i) u1[i] = z[m.up[b][i]];
#pragma unroll
  for (int i = 0; i < D2; ++i) u2[i] = z[m.down[b][i]];
  trafo_value<D2, D1>(trafo_at(th, o + T1, D2, D1, m), u2, s2, nullptr);
  if (m.variant == kJacEq1) {
#pragma unroll
    for (int i = 0; i < D1; ++i) u1[i] += s2[i];
  } else {
#pragma unroll
    for (int i = 0; i < D1; ++i) { u1[i] *= exp(s2[i]); lj += s2[i]; }
    if (m.variant == kAddS) {
#pragma unroll
      for (int i = 0; i < D1; ++i) u1[i] += s2[i];
    } else if (m.variant == kDifferentAdd) {
      double t2[D1];
      trafo_value<D2, D1>(trafo_at(th, o + 2 * T1 + T2, D2, D1, m), u2, t2, nullptr);
#pragma unroll
      for (int i = 0; i < D1; ++i) u1[i] += t2[i];
    }
  }
  trafo_value<D1, D2>(trafo_at(th, o, D1, D2, m), u1, s1, nullptr);
  if (m.variant == kJacEq1) {
#pragma unroll
    for (int i = 0; i < D2; ++i) u2[i] += s1[i];
  } else {
#pragma unroll
    for (int i = 0; i < D2; ++i) { u2[i] *= exp(s1[i]); lj += s1[i]; }
    if (m.variant == kAddS) {
#pragma unroll
      for (int i = 0; i < D2; ++i) u2[i] += s1[i];
    } else if (m.variant == kDifferentAdd) {
      double t1[D2];
      trafo_value<D1, D2>(trafo_at(th, o + T1 + T2, D1, D2, m), u1, t1, nullptr);
#pragma unroll
      for (int i = 0; i < D2; ++i) u2[i] += t1[i];
    }
  }
#pragma unroll
  for (int i = 0; i < D1; ++i) z[m.up[b][i]] = u1[i];
#pragma unroll
  for (int i = 0; i < D2; ++i) z[m.down[b][i]] = u2[i];
  if (has_global_change(m)) {  // net.py:115-116: scale * result + offset, log-Jacobian + d log(scale)
    const double* g = th + m.block_off[b];
    const double sc = g[D];
#pragma unroll
    for (int a = 0; a < D; ++a) z[a] = fma(sc, z[a], g[a]);
    lj += D * log(sc);
  }
  return lj;
}

// inverse block; returns the inverse map's log-Jacobian (= -(sum s1 + sum s2)), net.py:120-153
template <int D>
VMC_HD double block_inverse_value(const FlowMeta& m, const double* th, int b, double* z) {
  constexpr int D1 = D / 2, D2 = D - D / 2;
  const int T1 = trafo_size(D1, D2, m), T2 = trafo_size(D2, D1, m);
  const int o = block_base(m, b);
  double v1[D1], v2[D2], s1[D2], s2[D1], lj = 0.0;
#pragma unroll
  for (int i = 0; i < D1; ++i) v1[i] = z[m.up[b][i]];
#pragma unroll
  for (int i = 0; i < D2; ++i) v2[i] = z[m.down[b][i]];
  trafo_value<D1, D2>(trafo_at(th, o, D1, D2, m), v1, s1, nullptr);
  if (m.variant == kJacEq1) {
#pragma unroll
    for (int i = 0; i < D2; ++i) v2[i] -= s1[i];
  } else {
    if (m.variant == kAddS) {
#pragma unroll
      for (int i = 0; i < D2; ++i) v2[i] -= s1[i];
    } else if (m.variant == kDifferentAdd) {
      double t1[D2];
      trafo_value<D1, D2>(trafo_at(th, o + T1 + T2, D1, D2, m), v1, t1, nullptr);
#pragma unroll
      for (int i = 0; i < D2; ++i) v2[i] -= t1[i];
    }
#pragma unroll
    for (int i = 0; i < D2; ++i) { v2[i] *= exp(-s1[i]); lj -= s1[i]; }
  }
  trafo_value<D2, D1>(trafo_at(th, o + T1, D2, D1, m), v2, s2, nullptr);
  if (m.variant == kJacEq1) {
#pragma unroll
    for (int i = 0; i < D1; ++i) v1[i] -= s2[i];
  } else {
    if (m.variant == kAddS) {
#pragma unroll
      for (int i = 0; i < D1; ++i) v1[i] -= s2[i];
    } else if (m.variant == kDifferentAdd) {
      double t2[D1];
      trafo_value<D2, D1>(trafo_at(th, o + 2 * T1 + T2, D2, D1, m), v2, t2, nullptr);
#pragma unroll
      for (int i = 0; i < D1; ++i) v1[i] -= t2[i];
    }
#pragma unroll
    for (int i = 0; i < D1; ++i) { v1[i] *= exp(-s2[i]); lj -= s2[i]; }
  }
#pragma unroll
  for (int i = 0; i < D1; ++i) z[m.up[b][i]] = v1[i];
#pragma unroll
  for (int i = 0; i < D2; ++i) z[m.down[b][i]] = v2[i];
  if (has_global_change(m)) {
    // net.py:149-150, kept as the reference has it: the affine step is undone AFTER the inverse coupling, i.e. this is the
    // inverse of the forward block only for scale = 1, offset = 0 (the forward block applies it after the coupling, too)
    const double* g = th + m.block_off[b];
    const double sc = g[D];
#pragma unroll
    for (int a = 0; a < D; ++a) z[a] = (z[a] - g[a]) / sc;
    lj -= D * log(sc);
  }
  return lj;
}

// log p(x), net.py:209-213
template <int D>
VMC_HD double logp_value(const FlowMeta& m, const double* th, const double* x) {
  double z[D], lj = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) z[i] = x[i];
  for (int b = 0; b < m.depth; ++b) lj += block_forward_value<D>(m, th, b, z);
#pragma unroll
  for (int i = 0; i < D; ++i) z[i] -= m.offset[i];
  return latent_logpdf<D>(m, th, z) + lj;
}

// latent z -> (x, logp), net.py:214-217 (evaluate=False, inv=True)
template <int D>
VMC_HD double sample_from_latent(const FlowMeta& m, const double* th, const double* zlat, double* x) {
  double y[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { y[i] = zlat[i] - m.offset[i]; x[i] = zlat[i]; }
  const double plat = latent_logpdf<D>(m, th, y);
  double lj = 0.0;
  for (int b = m.depth - 1; b >= 0; --b) lj += block_inverse_value<D>(m, th, b, x);
  return plat - lj;
}

// ------------------------------------------------------------------------------------------------
// Jet forward through one block (in place on the jets of z), accumulating the log-Jacobian jet.
template <int D, int NT>
VMC_HD void block_forward_jet(const FlowMeta& m, const double* th, int b, const double* w, Jet<NT>* z, Jet<NT>& lj) {
  constexpr int D1 = D / 2, D2 = D - D / 2;
  const int T1 = trafo_size(D1, D2, m), T2 = trafo_size(D2, D1, m);
  const int o = block_base(m, b);
  Jet<NT> u1[D1], u2[D2];
#pragma unroll
  for (int i = 0; i < D1; ++i) u1[i] = z[m.up[b][i]];
#pragma unroll
  for (int i = 0; i < D2; ++i) u2[i] = z[m.down[b][i]];
  {
    Jet<NT> s2[D1];
    trafo_jet<NT, D2, D1>(trafo_at(th, o + T1, D2, D1, m), u2, w, s2);
    if (m.variant == kJacEq1) {
#pragma unroll
      for (int i = 0; i < D1; ++i) jet_add(u1[i], s2[i]);
    } else {
#pragma unroll
      for (int i = 0; i < D1; ++i) { jet_mulexp(u1[i], s2[i], w); jet_add(lj, s2[i]); }
      if (m.variant == kAddS) {
#pragma unroll
        for (int i = 0; i < D1; ++i) jet_add(u1[i], s2[i]);
      } else if (m.variant == kDifferentAdd) {
        trafo_jet<NT, D2, D1>(trafo_at(th, o + 2 * T1 + T2, D2, D1, m), u2, w, s2);
#pragma unroll
        for (int i = 0; i < D1; ++i) jet_add(u1[i], s2[i]);
      }
    }
  }
  {
    Jet<NT> s1[D2];
    trafo_jet<NT, D1, D2>(trafo_at(th, o, D1, D2, m), u1, w, s1);
    if (m.variant == kJacEq1) {
#pragma unroll
      for (int i = 0; i < D2; ++i) jet_add(u2[i], s1[i]);
    } else {
#pragma unroll
      for (int i = 0; i < D2; ++i) { jet_mulexp(u2[i], s1[i], w); jet_add(lj, s1[i]); }
      if (m.variant == kAddS) {
#pragma unroll
        for (int i = 0; i < D2; ++i) jet_add(u2[i], s1[i]);
      } else if (m.variant == kDifferentAdd) {
        trafo_jet<NT, D1, D2>(trafo_at(th, o + T1 + T2, D1, D2, m), u1, w, s1);
#pragma unroll
        for (int i = 0; i < D2; ++i) jet_add(u2[i], s1[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < D1; ++i) z[m.up[b][i]] = u1[i];
#pragma unroll
  for (int i = 0; i < D2; ++i) z[m.down[b][i]] = u2[i];
  if (has_global_change(m)) {
    const double* g = th + m.block_off[b];
    const double sc = g[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      z[a].v = fma(sc, z[a].v, g[a]);
      z[a].l *= sc;
#pragma unroll
      for (int k = 0; k < NT; ++k) z[a].g[k] *= sc;
    }
    lj.v += D * log(sc);
  }
}

// Result of the forward jet pass
template <int D, int NT = D>
struct JetResult {
  double logp;      // log p(x)
  double dir[NT];   // directional derivatives of log p along the tangents (grad_x if tangents = I)
  double lap;       // sum_k w_k d^2 logp / dv_k^2
  double zfin[D];   // flow output z = INN(x) (needed to start the reverse sweep)
};

// forward jets of log p.  tang = nullptr means identity tangents (NT == D); else row-major [NT][D], row k = v_k.
template <int D, int NT = D>
VMC_HD void logp_jet(const FlowMeta& m, const double* th, const double* x, const double* tang,
                     const double* w, JetResult<D, NT>& res) {
  Jet<NT> z[D], lj;
  jet_const(lj, 0.0);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    jet_const(z[i], x[i]);
#pragma unroll
    for (int k = 0; k < NT; ++k) z[i].g[k] = tang ? tang[k * D + i] : (k == i ? 1.0 : 0.0);
  }
  for (int b = 0; b < m.depth; ++b) block_forward_jet<D, NT>(m, th, b, w, z, lj);
  double L[D][D];
  build_L<D>(m, th, L);
  Jet<NT> wv[D];
  double sumLd = 0.0;
#pragma unroll
  for (int a = D - 1; a >= 0; --a) {
    res.zfin[a] = z[a].v;
    Jet<NT> s = z[a];
    s.v -= m.offset[a] + th[m.off_mu + a];
#pragma unroll
    for (int b2 = a + 1; b2 < D; ++b2) jet_axpy(s, -L[a][b2], wv[b2]);
    const double inv = 1.0 / L[a][a];
    wv[a].v = s.v * inv; wv[a].l = s.l * inv;
#pragma unroll
    for (int k = 0; k < NT; ++k) wv[a].g[k] = s.g[k] * inv;
    sumLd += th[m.off_Ldiag + a];
  }
  Jet<NT> q;
  jet_const(q, 0.0);
#pragma unroll
  for (int a = 0; a < D; ++a) {
    q.v += wv[a].v * wv[a].v;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NT; ++k) { q.g[k] += 2.0 * wv[a].v * wv[a].g[k]; s += w[k] * wv[a].g[k] * wv[a].g[k]; }
    q.l += 2.0 * (wv[a].v * wv[a].l + s);
  }
  double c, f1, f2;
  if (m.latent == kGauss) {
    c = -0.5 * (D * 1.8378770664093454835606594728112 + 2.0 * sumLd) - 0.5 * q.v;
    f1 = -0.5; f2 = 0.0;
  } else {
    const double nu = exp(th[m.off_dist]) + 1.0;
    c = lgamma(0.5 * (nu + D)) - lgamma(0.5 * nu) - 0.5 * D * log(nu * 3.14159265358979323846) -
        0.5 * (nu + D) * log(1.0 + q.v / nu);
    f1 = -0.5 * (nu + D) / (nu + q.v);
    f2 = 0.5 * (nu + D) / ((nu + q.v) * (nu + q.v));
  }
  res.logp = c + lj.v;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NT; ++k) { res.dir[k] = f1 * q.g[k] + lj.g[k]; s += w[k] * q.g[k] * q.g[k]; }
  res.lap = f1 * q.l + f2 * s + lj.l;
}

// ------------------------------------------------------------------------------------------------
// Reverse sweep: emits O row (flat order) given the flow output zfin; returns grad_x in gx (optional).
// Memory-free: block inputs are reconstructed from block outputs with the inverse coupling.
template <int D, class Emit>
VMC_HD void logp_reverse(const FlowMeta& m, const double* th, const double* zfin, Emit& em, double* gx) {
  constexpr int D1 = D / 2, D2 = D - D / 2;
  double L[D][D];
  build_L<D>(m, th, L);
  double wv[D], r[D], q = 0.0;
#pragma unroll
  for (int a = D - 1; a >= 0; --a) {
    double s = zfin[a] - m.offset[a] - th[m.off_mu + a];
#pragma unroll
    for (int b = a + 1; b < D; ++b) s -= L[a][b] * wv[b];
    wv[a] = s / L[a][a];
    q += wv[a] * wv[a];
  }
#pragma unroll
  for (int a = 0; a < D; ++a) {  // r = L^-T w
    double s = wv[a];
#pragma unroll
    for (int b = 0; b < a; ++b) s -= L[b][a] * r[b];
    r[a] = s / L[a][a];
  }
  double f1 = -0.5, nu = 0.0;
  if (m.latent == kStudentT) { nu = exp(th[m.off_dist]) + 1.0; f1 = -0.5 * (nu + D) / (nu + q); }
  // latent segment: L, L_diag, dist_params, mu (sorted keys, uppercase first)
  em.seek(0);
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a + 1; b < D; ++b) em.put(-2.0 * f1 * r[a] * wv[b]);
#pragma unroll
  for (int a = 0; a < D; ++a)
    em.put(-2.0 * f1 * r[a] * wv[a] * L[a][a] - (m.latent == kGauss ? 1.0 : 0.0));
  if (m.latent == kStudentT) {
    const double dnu = 0.5 * digamma_pos(0.5 * (nu + D)) - 0.5 * digamma_pos(0.5 * nu) - 0.5 * D / nu -
                       0.5 * log(1.0 + q / nu) + 0.5 * (nu + D) * q / (nu * (nu + q));
    em.put((nu - 1.0) * dnu);
  }
#pragma unroll
  for (int a = 0; a < D; ++a) em.put(-2.0 * f1 * r[a]);

  double z[D], dz[D];
#pragma unroll
  for (int a = 0; a < D; ++a) { z[a] = zfin[a]; dz[a] = 2.0 * f1 * r[a]; }

  const int T1 = trafo_size(D1, D2, m), T2 = trafo_size(D2, D1, m);
  double hid[kMaxHidden], dh[kMaxHidden];
  for (int b = m.depth - 1; b >= 0; --b) {
    const int o = block_base(m, b);
    if (has_global_change(m)) {
      // block output = scale * r + offset with r the coupling's output: d/d offset_a = dz_a, d/d scale = sum_a dz_a r_a + d / scale
      const double* g = th + m.block_off[b];
      const double sc = g[D];
      double dsc = D / sc;
      em.seek(m.block_off[b]);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        em.put(dz[a]);
        z[a] = (z[a] - g[a]) / sc;
        dsc = fma(dz[a], z[a], dsc);
        dz[a] *= sc;
      }
      em.put(dsc);
    }
    const Trafo ts1 = trafo_at(th, o, D1, D2, m), ts2 = trafo_at(th, o + T1, D2, D1, m);
    const Trafo tt1 = trafo_at(th, o + T1 + T2, D1, D2, m), tt2 = trafo_at(th, o + 2 * T1 + T2, D2, D1, m);
    double v1[D1], v2[D2], dv1[D1], dv2[D2];
#pragma unroll
    for (int i = 0; i < D1; ++i) { v1[i] = z[m.up[b][i]]; dv1[i] = dz[m.up[b][i]]; }
#pragma unroll
    for (int i = 0; i < D2; ++i) { v2[i] = z[m.down[b][i]]; dv2[i] = dz[m.down[b][i]]; }
    // ---- second half of the block: v2 = u2 * exp(s1(v1)) [+ t1(v1) | + s1]
    double s1[D2], t1[D2], u2[D2], ds1[D2], du2[D2];
    trafo_value<D1, D2>(ts1, v1, s1, hid);
    if (m.variant == kJacEq1) {
#pragma unroll
      for (int i = 0; i < D2; ++i) { u2[i] = v2[i] - s1[i]; ds1[i] = dv2[i]; du2[i] = dv2[i]; }
    } else {
      if (m.variant == kDifferentAdd) trafo_value<D1, D2>(tt1, v1, t1, nullptr);
#pragma unroll
      for (int i = 0; i < D2; ++i) {
        const double e = exp(s1[i]);
        const double sub = m.variant == kDifferentAdd ? t1[i] : (m.variant == kAddS ? s1[i] : 0.0);
        u2[i] = (v2[i] - sub) / e;
        ds1[i] = dv2[i] * (u2[i] * e + (m.variant == kAddS ? 1.0 : 0.0)) + 1.0;
        du2[i] = dv2[i] * e;
      }
    }
    em.seek(o);
    trafo_reverse<D1, D2, true, true>(ts1, v1, s1, ds1, hid, dh, dv1, em);
    if (m.variant == kDifferentAdd) {  // t1's input gradient is needed before s2; its parameters come later
      trafo_value<D1, D2>(tt1, v1, t1, hid);
      trafo_reverse<D1, D2, false, true>(tt1, v1, t1, dv2, hid, dh, dv1, em);
    }
    // ---- first half: v1 = u1 * exp(s2(u2)) [+ t2(u2) | + s2]
    double s2[D1], t2[D1], u1[D1], ds2[D1], du1[D1];
    trafo_value<D2, D1>(ts2, u2, s2, hid);
    if (m.variant == kJacEq1) {
#pragma unroll
      for (int i = 0; i < D1; ++i) { u1[i] = v1[i] - s2[i]; ds2[i] = dv1[i]; du1[i] = dv1[i]; }
    } else {
      if (m.variant == kDifferentAdd) trafo_value<D2, D1>(tt2, u2, t2, nullptr);
#pragma unroll
      for (int i = 0; i < D1; ++i) {
        const double e = exp(s2[i]);
        const double sub = m.variant == kDifferentAdd ? t2[i] : (m.variant == kAddS ? s2[i] : 0.0);
        u1[i] = (v1[i] - sub) / e;
        ds2[i] = dv1[i] * (u1[i] * e + (m.variant == kAddS ? 1.0 : 0.0)) + 1.0;
        du1[i] = dv1[i] * e;
      }
    }
    trafo_reverse<D2, D1, true, true>(ts2, u2, s2, ds2, hid, dh, du2, em);
    if (m.variant == kDifferentAdd) {
      double dummy[D1];
      trafo_value<D1, D2>(tt1, v1, t1, hid);
      trafo_reverse<D1, D2, true, false>(tt1, v1, t1, dv2, hid, dh, dummy, em);
      trafo_value<D2, D1>(tt2, u2, t2, hid);
      trafo_reverse<D2, D1, true, true>(tt2, u2, t2, dv1, hid, dh, du2, em);
    }
#pragma unroll
    for (int i = 0; i < D1; ++i) { z[m.up[b][i]] = u1[i]; dz[m.up[b][i]] = du1[i]; }
#pragma unroll
    for (int i = 0; i < D2; ++i) { z[m.down[b][i]] = u2[i]; dz[m.down[b][i]] = du2[i]; }
  }
  em.finish();
  if (gx) {
#pragma unroll
    for (int a = 0; a < D; ++a) gx[a] = dz[a];
  }
}

// ------------------------------------------------------------------------------------------------
// Laplacian weights per tangent and the local term E_loc = d_t log p (evolutionEq.py:84-119).
template <int D>
VMC_HD void equation_weights(const EqParams& e, double* w) {
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (e.mode == kAdvectionHamiltonianWDiss) w[k] = (k & 1) ? 1.0 : 0.0;
    else if (e.mode == kAdvectionHamiltonian || e.mode == kAdvectionPaper) w[k] = 0.0;
    else w[k] = 1.0;
  }
}

template <int D>
VMC_HD void velocity_hamiltonian(const EqParams& e, const double* x, double* v) {
  // evolutionEq.py:30-45 (uncoupled): v = J grad H on interleaved (x,p): (p/m, -m w^2 x - 4 lam x^3)
#pragma unroll
  for (int i = 0; i + 1 < D; i += 2) {
    v[i] = x[i + 1] / e.m;
    v[i + 1] = -(e.m * e.omega * e.omega * x[i] + 4.0 * e.lam * x[i] * x[i] * x[i]);
  }
  if (D & 1) v[D - 1] = 0.0;
}

template <int D>
VMC_HD double local_term(const EqParams& e, const double* x, const JetResult<D>& r) {
  const double pi = 3.14159265358979323846;
  double g2 = 0.0, gs = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) { g2 += r.dir[k] * r.dir[k]; gs += r.dir[k]; }
  switch (e.mode) {
    case kDiffusion: return e.D * (g2 + r.lap);
    case kDiffusionDrift: return e.D * (g2 + r.lap) + e.mu * gs;
    case kDiffusionAniso: return g2 + r.lap;  // tangents = rows of A, D = A^T A
    case kAdvectionHamiltonian:
    case kAdvectionHamiltonianWDiss: {
      double v[D], adv = 0.0;
      velocity_hamiltonian<D>(e, x, v);
#pragma unroll
      for (int k = 0; k < D; ++k) adv -= r.dir[k] * v[k];
      if (e.mode == kAdvectionHamiltonian) return adv;
      double godd2 = 0.0, damp = 0.0;
#pragma unroll
      for (int k = 1; k < D; k += 2) { godd2 += r.dir[k] * r.dir[k]; damp += x[k] * r.dir[k]; }
      return adv + e.m * e.gamma * e.T * (godd2 + r.lap) + e.gamma * damp;
    }
    case kAdvectionPaper: {
      if (D < 2) return 0.0;
      const double c = cos(pi * e.t / e.T);
      const double sx = sin(pi * x[0]), sy = sin(pi * x[D > 1 ? 1 : 0]);
      const double vx = -sx * sx * sin(2.0 * pi * x[D > 1 ? 1 : 0]) * c;
      const double vy = sy * sy * sin(2.0 * pi * x[0]) * c;
      return -(r.dir[0] * vx + r.dir[D > 1 ? 1 : 0] * vy);
    }
  }
  return 0.0;
}

// Full Hessian of log p wrt x (var_state.py:32,66-67) by polarisation of single-tangent jets:
// H_ii = d^2/de_i^2, H_ij = (d^2/d(e_i+e_j)^2 - H_ii - H_jj)/2.  API completeness only; the fused
// local-terms path never forms H.
template <int D>
VMC_HD void logp_hessian(const FlowMeta& m, const double* th, const double* x, double* H) {
  const double w1[1] = {1.0};
  double tang[D];
  for (int i = 0; i < D; ++i) {
    for (int k = 0; k < D; ++k) tang[k] = (k == i) ? 1.0 : 0.0;
    JetResult<D, 1> r;
    logp_jet<D, 1>(m, th, x, tang, w1, r);
    H[i * D + i] = r.lap;
  }
  for (int i = 0; i < D; ++i)
    for (int j = i + 1; j < D; ++j) {
      for (int k = 0; k < D; ++k) tang[k] = (k == i || k == j) ? 1.0 : 0.0;
      JetResult<D, 1> r;
      logp_jet<D, 1>(m, th, x, tang, w1, r);
      const double v = 0.5 * (r.lap - H[i * D + i] - H[j * D + j]);
      H[i * D + j] = v; H[j * D + i] = v;
    }
}

// Trivial emitter for host tests / row-at-a-time use: writes straight into a row.
struct RowEmit {
  double* row;
  int p;
  VMC_HD void seek(int q) { p = q; }
  VMC_HD void put(double v) { row[p++] = v; }
  VMC_HD void finish() {}
};

}  // namespace vmc
