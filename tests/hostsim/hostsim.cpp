// TEST TOOLING ONLY.  Instantiates the __host__ __device__ per-sample math of
// vmc_pde_b200/csrc/flow_core.cuh on the CPU so that the hand-derived jets and reverse sweep can be
// checked against the torch-autograd oracle without a GPU.  Not linked into libvmcpde.so; the product
// path has no CPU fallback.
#include <cstring>
#include <string>
#include "../../vmc_pde_b200/csrc/flow_meta.hpp"

using namespace vmc;

template <int D>
static int local_terms_t(const FlowMeta& m, const double* th, const double* x, long n, const vmcpde_equation* eq,
                         const double* tang, double* eloc, double* logp, double* grad, double* lap, double* O,
                         long ldo, double* gx_rev) {
  EqParams e{eq->mode, eq->D, eq->mu, eq->m, eq->omega, eq->lam, eq->T, eq->gamma, eq->t};
  double w[D];
  equation_weights<D>(e, w);
  for (long i = 0; i < n; ++i) {
    JetResult<D> r;
    logp_jet<D>(m, th, x + i * D, tang, w, r);
    if (eloc) eloc[i] = local_term<D>(e, x + i * D, r);
    if (logp) logp[i] = r.logp;
    if (lap) lap[i] = r.lap;
    if (grad) for (int k = 0; k < D; ++k) grad[i * D + k] = r.dir[k];
    if (O) {
      RowEmit em{O + i * ldo, 0};
      logp_reverse<D>(m, th, r.zfin, em, gx_rev ? gx_rev + i * D : nullptr);
    }
  }
  return 0;
}

extern "C" int hostsim_num_params(const vmcpde_flow_config* c) {
  FlowMeta m; std::string err;
  if (make_flow_meta(c, &m, &err)) return -1;
  return m.P;
}

extern "C" int hostsim_local_terms(const vmcpde_flow_config* c, const double* th, const double* x, long n,
                                   const vmcpde_equation* eq, double* eloc, double* logp, double* grad, double* lap,
                                   double* O, long ldo, double* gx_rev) {
  FlowMeta m; std::string err;
  int rc = make_flow_meta(c, &m, &err);
  if (rc) return rc;
  VMC_DISPATCH_DIM(m.d, return local_terms_t<D>(m, th, x, n, eq, eq->tangents, eloc, logp, grad, lap, O, ldo, gx_rev));
  return 0;
}

template <int D>
static int sample_t(const FlowMeta& m, const double* th, const double* z, long n, double* x, double* logp) {
  for (long i = 0; i < n; ++i) logp[i] = sample_from_latent<D>(m, th, z + i * D, x + i * D);
  return 0;
}
extern "C" int hostsim_sample_from_latent(const vmcpde_flow_config* c, const double* th, const double* z, long n,
                                          double* x, double* logp) {
  FlowMeta m; std::string err;
  int rc = make_flow_meta(c, &m, &err);
  if (rc) return rc;
  VMC_DISPATCH_DIM(m.d, return sample_t<D>(m, th, z, n, x, logp));
  return 0;
}

template <int D>
static int logp_t(const FlowMeta& m, const double* th, const double* x, long n, double* logp) {
  for (long i = 0; i < n; ++i) logp[i] = logp_value<D>(m, th, x + i * D);
  return 0;
}
extern "C" int hostsim_logp(const vmcpde_flow_config* c, const double* th, const double* x, long n, double* logp) {
  FlowMeta m; std::string err;
  int rc = make_flow_meta(c, &m, &err);
  if (rc) return rc;
  VMC_DISPATCH_DIM(m.d, return logp_t<D>(m, th, x, n, logp));
  return 0;
}

template <int D>
static int hess_t(const FlowMeta& m, const double* th, const double* x, long n, double* H) {
  for (long i = 0; i < n; ++i) logp_hessian<D>(m, th, x + i * D, H + i * D * D);
  return 0;
}
extern "C" int hostsim_hessian(const vmcpde_flow_config* c, const double* th, const double* x, long n, double* H) {
  FlowMeta m; std::string err;
  int rc = make_flow_meta(c, &m, &err);
  if (rc) return rc;
  VMC_DISPATCH_DIM(m.d, return hess_t<D>(m, th, x, n, H));
  return 0;
}

// ---- host driver of the tridiagonal divide & conquer (same scalar core as eigh.cu) ------------------
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "../../vmc_pde_b200/csrc/dc_core.cuh"
extern "C" int hostsim_dc_eigh(int n, const double* d, const double* e, double* lam, double* QT) {
  std::vector<double> l(n), Qa((size_t)n * n, 0.0), Qb((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    l[i] = d[i] - (i > 0 ? fabs(e[i - 1]) : 0.0) - (i + 1 < n ? fabs(e[i]) : 0.0);
    Qa[(size_t)i * n + i] = 1.0;
  }
  const int D = dc_tree_depth(n);
  std::vector<double> z(n), dmod(n), dl(n), w(n), w2(n), dfv(n), tau(n), what(n), vals(n), lnew(n);
  std::vector<int> order(n), nd(n), dfi(n), org(n), pos(n);
  std::vector<DcRot> rots(n);
  double* Q = Qa.data();
  double* Qn = Qb.data();
  for (int depth = D - 1; depth >= 0; --depth) {
    for (int idx = 0; idx < (1 << depth); ++idx) {
      int lo, hi;
      dc_node_range(n, depth, idx, lo, hi);
      const int nm = hi - lo, mid = lo + nm / 2, n1 = mid - lo;
      if (nm <= 1 || n1 == 0) {  // nothing to merge: carry over
        for (int r = lo; r < hi; ++r) for (int c = lo; c < hi; ++c) Qn[(size_t)r * n + c] = Q[(size_t)r * n + c];
        for (int r = lo; r < hi; ++r) lnew[r] = l[r];
        continue;
      }
      const double rs = e[mid - 1];
      for (int c = 0; c < nm; ++c) z[c] = c < n1 ? Q[(size_t)(lo + c) * n + mid - 1] : (rs < 0 ? -1.0 : 1.0) * Q[(size_t)(lo + c) * n + mid];
      int k, nrot; double rho;
      dc_deflate(nm, n1, rs, &l[lo], z.data(), dmod.data(), order.data(), dl.data(), w.data(), nd.data(), dfv.data(), dfi.data(), rots.data(), &k, &nrot, &rho);
      for (int r = 0; r < nrot; ++r)
        for (int c = lo; c < hi; ++c) {
          double& x = Q[(size_t)(lo + rots[r].a) * n + c]; double& y = Q[(size_t)(lo + rots[r].b) * n + c];
          const double xn = rots[r].c * x + rots[r].s * y, yn = rots[r].c * y - rots[r].s * x;
          x = xn; y = yn;
        }
      for (int i = 0; i < k; ++i) w2[i] = w[i] * w[i];
      SerialSums ss{dl.data(), w2.data(), k};
      for (int j = 0; j < k; ++j) { secular_root(k, j, dl.data(), w2.data(), rho, ss, &org[j], &tau[j]); vals[j] = dl[org[j]] + tau[j]; }
      for (int m = 0; m < nm - k; ++m) vals[k + m] = dfv[m];
      if (getenv("HOSTSIM_DC_DEBUG") && nm > 200) {
        printf("merge nm=%d k=%d nrot=%d rho=%.3e\n", nm, k, nrot, rho);
        for (int j = 0; j < k; ++j) {
          SecularSums q = ss(org[j], tau[j], j == k - 1 ? k - 2 : j);
          double f = 1.0 / rho + q.psi + q.phi;
          double gap = j + 1 < k ? dl[j + 1] - dl[j] : 0.0;
          if (fabs(f) > 1e-10 * q.sabs || j >= k - 6) printf("  root %d org %d tau %.6e gap %.3e f %.3e sabs %.3e fprime %.3e w2[j] %.3e w2[j+1] %.3e\n", j, org[j], tau[j], gap, f, q.sabs, q.dpsi + q.dphi, w2[j], j + 1 < k ? w2[j + 1] : 0.0);
        }
      }
      for (int i = 0; i < nm; ++i) { int r = 0; for (int q = 0; q < nm; ++q) r += (vals[q] < vals[i]) || (vals[q] == vals[i] && q < i); pos[i] = r; }
      for (int i = 0; i < k; ++i) what[i] = dc_lowner_w(k, i, dl.data(), org.data(), tau.data(), w[i]);
      for (int j = 0; j < k; ++j) {
        double nrm = 0.0;
        for (int i = 0; i < k; ++i) { const double u = what[i] / dc_delta(dl.data(), i, org[j], tau[j]); nrm += u * u; }
        nrm = sqrt(nrm);
        double* out = &Qn[(size_t)(lo + pos[j]) * n];
        for (int c = lo; c < hi; ++c) out[c] = 0.0;
        for (int i = 0; i < k; ++i) {
          const double u = what[i] / dc_delta(dl.data(), i, org[j], tau[j]) / nrm;
          const double* src = &Q[(size_t)(lo + nd[i]) * n];
          for (int c = lo; c < hi; ++c) out[c] += u * src[c];
        }
        lnew[lo + pos[j]] = vals[j];
      }
      for (int m = 0; m < nm - k; ++m) {
        const double* src = &Q[(size_t)(lo + dfi[m]) * n];
        double* out = &Qn[(size_t)(lo + pos[k + m]) * n];
        for (int c = lo; c < hi; ++c) out[c] = src[c];
        lnew[lo + pos[k + m]] = vals[k + m];
      }
    }
    std::swap(Q, Qn);
    l = lnew;
    if (getenv("HOSTSIM_DC_DEBUG")) {
      for (int idx = 0; idx < (1 << depth); ++idx) {
        int lo, hi; dc_node_range(n, depth, idx, lo, hi);
        const int nm = hi - lo; if (nm <= 1) continue;
        double worst = 0; int wj = -1;
        for (int j = lo; j < hi; ++j) {   // residual of eigenpair j of the node matrix
          for (int r = lo; r < hi; ++r) {
            double dd = d[r] - ((r == lo && lo > 0) ? fabs(e[lo - 1]) : 0.0) - ((r == hi - 1 && hi < n) ? fabs(e[hi - 1]) : 0.0);
            double v = dd * Q[(size_t)j * n + r];
            if (r > lo) v += e[r - 1] * Q[(size_t)j * n + r - 1];
            if (r + 1 < hi) v += e[r] * Q[(size_t)j * n + r + 1];
            v -= l[j] * Q[(size_t)j * n + r];
            if (fabs(v) > worst) { worst = fabs(v); wj = j - lo; }
          }
        }
        if (worst > 2e-14) printf("depth %d node %d [%d,%d) worst resid %.2e at local eig %d of %d\n", depth, idx, lo, hi, worst, wj, nm);
      }
    }
  }
  for (int i = 0; i < n; ++i) lam[i] = l[i];
  for (size_t i = 0; i < (size_t)n * n; ++i) QT[i] = Q[i];
  return 0;
}
extern "C" int hostsim_secular(int k, const double* dl, const double* w2, double rho, int* origin, double* tau) {
  SerialSums ss{dl, w2, k};
  for (int j = 0; j < k; ++j) secular_root(k, j, dl, w2, rho, ss, &origin[j], &tau[j]);
  return 0;
}
