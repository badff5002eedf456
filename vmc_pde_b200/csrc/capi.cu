// C-ABI glue: error state, ansatz handle, dispatch of the per-dimension flow kernels.
#include <string>
#include "common.cuh"
#include "flow_kernels.cuh"
#include "flow_meta.hpp"
#include "rng.cuh"

namespace vmc {
std::string& last_error_ref() {
  static thread_local std::string e;
  return e;
}
int set_error(int code, const std::string& msg) {
  last_error_ref() = msg;
  return code;
}

#define VMC_DECL2(Dv, MLv)                                                                                                  \
  extern template int launch_sample<Dv, MLv>(const FlowMeta&, const double*, uint32_t, uint32_t, long long, long long,     \
                                        long long, const double*, double*, double*, double*, cudaStream_t);          \
  extern template int launch_logp<Dv, MLv>(const FlowMeta&, const double*, const double*, long long, double*, cudaStream_t); \
  extern template int launch_local_terms<Dv, MLv>(const FlowMeta&, const double*, const double*, long long, const EqParams&, \
                                             const double*, double*, double*, double*, double*, double*, long long,   \
                                             cudaStream_t);                                                           \
  extern template int launch_transform<Dv, MLv>(const FlowMeta&, const double*, const double*, long long, int, double*,    \
                                           double*, double*, cudaStream_t);                                           \
  extern template int launch_hessian<Dv, MLv>(const FlowMeta&, const double*, const double*, long long, double*, cudaStream_t);
#define VMC_DECL(Dv) VMC_DECL2(Dv, 0) VMC_DECL2(Dv, 1)
VMC_FOR_EACH_DIM(VMC_DECL)

__global__ void normal_kernel(uint32_t k0, uint32_t k1, long long first, long long n, unsigned long long total,
                              double* __restrict__ out, int uniform) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t bits = random_bits64(k0, k1, (unsigned long long)(first + i), total);
  out[i] = uniform ? bits_to_unit(bits) : normal_from_bits(bits);
}
}  // namespace vmc

using namespace vmc;

extern "C" __attribute__((visibility("default"))) const char* vmcpde_last_error(void) { return last_error_ref().c_str(); }
extern "C" __attribute__((visibility("default"))) int vmcpde_version(void) { return VMCPDE_VERSION; }
extern "C" __attribute__((visibility("default"))) int32_t vmcpde_padded_params(int32_t p) { return p <= 0 ? 128 : ((p + 127) / 128) * 128; }

extern "C" __attribute__((visibility("default"))) int vmcpde_flow_create(const vmcpde_flow_config* cfg, vmcpde_flow** out) {
  VMC_REQUIRE(cfg && out, "vmcpde_flow_create: null argument");
  vmcpde_flow* f = new vmcpde_flow();
  std::string err;
  int rc = make_flow_meta(cfg, &f->meta, &err);
  if (rc) { delete f; return set_error(rc, "vmcpde_flow_create: " + err); }
  bool built = false;
#define VMC_CHECK_DIM(Dv) if (cfg->dim == Dv) built = true;
  VMC_FOR_EACH_DIM(VMC_CHECK_DIM)
  if (!built) { delete f; return set_error(VMCPDE_EUNSUPPORTED, "vmcpde_flow_create: dimension not built (2-6, 8, 10, 12)"); }
  *out = f;
  return 0;
}
extern "C" __attribute__((visibility("default"))) void vmcpde_flow_destroy(vmcpde_flow* f) { delete f; }
extern "C" __attribute__((visibility("default"))) int32_t vmcpde_flow_num_params(const vmcpde_flow* f) { return f ? f->meta.P : -1; }
extern "C" __attribute__((visibility("default"))) int vmcpde_flow_param_offsets(const vmcpde_flow* f, int32_t* out) {
  VMC_REQUIRE(f && out, "vmcpde_flow_param_offsets: null argument");
  out[0] = f->meta.off_L; out[1] = f->meta.off_Ldiag; out[2] = f->meta.off_dist; out[3] = f->meta.off_mu;
  for (int b = 0; b < f->meta.depth; ++b) out[4 + b] = f->meta.block_off[b];
  return 0;
}

#define VMC_CASE(Dv, call) case Dv: { constexpr int D = Dv; return call; }

extern "C" __attribute__((visibility("default"))) int vmcpde_sample(const vmcpde_flow* f, const double* theta, uint32_t key0, uint32_t key1, int64_t first,
                             int64_t n, int64_t n_total, const double* chi2, double* x, double* logp, double* z_out,
                             vmcpde_stream stream) {
  VMC_REQUIRE(f && theta && x && logp, "vmcpde_sample: null pointer");
  VMC_REQUIRE(first >= 0 && n >= 0 && first + n <= n_total, "vmcpde_sample: [first, first+n) must lie in [0, n_total)");
  VMC_REQUIRE((unsigned long long)n_total * f->meta.d * 2ull <= 0x100000000ull,
              "vmcpde_sample: n_total * dim * 2 exceeds the 32-bit threefry counter range");
  VMC_REQUIRE(f->meta.latent != kStudentT || chi2, "vmcpde_sample: Student_t needs chi2 variates");
  const FlowMeta& m = f->meta;
  cudaStream_t s = (cudaStream_t)stream;
  switch (m.d) {
#define X(Dv) VMC_CASE(Dv, ((m.nl > 1 || m.gc) ? launch_sample<D, 1>(m, theta, key0, key1, first, n, n_total, chi2, x, logp, z_out, s) : launch_sample<D, 0>(m, theta, key0, key1, first, n, n_total, chi2, x, logp, z_out, s)))
    VMC_FOR_EACH_DIM(X)
#undef X
  }
  return set_error(VMCPDE_EUNSUPPORTED, "dimension not built");
}

static int launch_rng(uint32_t k0, uint32_t k1, int64_t first, int64_t n, int64_t total, double* out, void* stream, int uni) {
  VMC_REQUIRE(out && first >= 0 && n >= 0 && first + n <= total, "rng: bad range");
  VMC_REQUIRE((unsigned long long)total * 2ull <= 0x100000000ull, "rng: count exceeds the 32-bit threefry counter range");
  if (n == 0) return 0;
  normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(k0, k1, first, n, (unsigned long long)total, out, uni);
  VMC_LAUNCH_CHECK("normal_kernel");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int vmcpde_normal(uint32_t key0, uint32_t key1, int64_t first, int64_t n, int64_t count_total, double* out,
                             vmcpde_stream stream) {
  return launch_rng(key0, key1, first, n, count_total, out, stream, 0);
}
extern "C" __attribute__((visibility("default"))) int vmcpde_uniform(uint32_t key0, uint32_t key1, int64_t first, int64_t n, int64_t count_total, double* out,
                              vmcpde_stream stream) {
  return launch_rng(key0, key1, first, n, count_total, out, stream, 1);
}

extern "C" __attribute__((visibility("default"))) int vmcpde_logp(const vmcpde_flow* f, const double* theta, const double* x, int64_t n, double* logp,
                           vmcpde_stream stream) {
  VMC_REQUIRE(f && theta && x && logp, "vmcpde_logp: null pointer");
  const FlowMeta& m = f->meta;
  cudaStream_t s = (cudaStream_t)stream;
  switch (m.d) {
#define X(Dv) VMC_CASE(Dv, ((m.nl > 1 || m.gc) ? launch_logp<D, 1>(m, theta, x, n, logp, s) : launch_logp<D, 0>(m, theta, x, n, logp, s)))
    VMC_FOR_EACH_DIM(X)
#undef X
  }
  return set_error(VMCPDE_EUNSUPPORTED, "dimension not built");
}

extern "C" __attribute__((visibility("default"))) int vmcpde_local_terms(const vmcpde_flow* f, const double* theta, const double* x, int64_t n,
                                  const vmcpde_equation* eq, double* eloc, double* logp, double* grad, double* lap,
                                  double* O, int64_t ldo, vmcpde_stream stream) {
  VMC_REQUIRE(f && theta && x && eq, "vmcpde_local_terms: null pointer");
  VMC_REQUIRE(eq->mode >= 0 && eq->mode <= 5, "vmcpde_local_terms: unknown equation");
  VMC_REQUIRE(eq->mode != VMCPDE_DIFFUSION_ANISOTROPIC || eq->tangents, "vmcpde_local_terms: anisotropic diffusion needs tangents");
  VMC_REQUIRE(eq->mode != VMCPDE_ADVECTION_PAPER || f->meta.d == 2, "vmcpde_local_terms: advection_paper is two-dimensional");
  VMC_REQUIRE(!O || ldo >= f->meta.P, "vmcpde_local_terms: ldo < num_params");
  const FlowMeta& m = f->meta;
  EqParams e{eq->mode, eq->D, eq->mu, eq->m, eq->omega, eq->lam, eq->T, eq->gamma, eq->t};
  cudaStream_t s = (cudaStream_t)stream;
  switch (m.d) {
#define X(Dv) VMC_CASE(Dv, ((m.nl > 1 || m.gc) ? launch_local_terms<D, 1>(m, theta, x, n, e, eq->tangents, eloc, logp, grad, lap, O, ldo, s) : launch_local_terms<D, 0>(m, theta, x, n, e, eq->tangents, eloc, logp, grad, lap, O, ldo, s)))
    VMC_FOR_EACH_DIM(X)
#undef X
  }
  return set_error(VMCPDE_EUNSUPPORTED, "dimension not built");
}

extern "C" __attribute__((visibility("default"))) int vmcpde_hessian(const vmcpde_flow* f, const double* theta, const double* x, int64_t n, double* H,
                              vmcpde_stream stream) {
  VMC_REQUIRE(f && theta && x && H, "vmcpde_hessian: null pointer");
  const FlowMeta& m = f->meta;
  cudaStream_t s = (cudaStream_t)stream;
  switch (m.d) {
#define X(Dv) VMC_CASE(Dv, ((m.nl > 1 || m.gc) ? launch_hessian<D, 1>(m, theta, x, n, H, s) : launch_hessian<D, 0>(m, theta, x, n, H, s)))
    VMC_FOR_EACH_DIM(X)
#undef X
  }
  return set_error(VMCPDE_EUNSUPPORTED, "dimension not built");
}

extern "C" __attribute__((visibility("default"))) int vmcpde_flow_transform(const vmcpde_flow* f, const double* theta, const double* x,
                                                                           int64_t n, int32_t inverse, double* y, double* logjac,
                                                                           double* latent_logpdf_of_input, vmcpde_stream stream) {
  VMC_REQUIRE(f && theta && x && y, "vmcpde_flow_transform: null pointer");
  const FlowMeta& m = f->meta;
  cudaStream_t s = (cudaStream_t)stream;
  switch (m.d) {
#define X(Dv) VMC_CASE(Dv, ((m.nl > 1 || m.gc) ? launch_transform<D, 1>(m, theta, x, n, inverse, y, logjac, latent_logpdf_of_input, s) : launch_transform<D, 0>(m, theta, x, n, inverse, y, logjac, latent_logpdf_of_input, s)))
    VMC_FOR_EACH_DIM(X)
#undef X
  }
  return set_error(VMCPDE_EUNSUPPORTED, "dimension not built");
}
