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
