// Launchers of the per-dimension flow kernels (defined in flow_kernels_dim.cu, one object per D).
#pragma once
#include "common.cuh"
#include "flow_core.cuh"

struct vmcpde_flow {
  vmc::FlowMeta meta;
};

namespace vmc {

template <int D, int ML>
int launch_sample(const FlowMeta& m, const double* theta, uint32_t k0, uint32_t k1, long long first, long long n,
                  long long n_total, const double* chi2, double* x, double* logp, double* zout, cudaStream_t s);
template <int D, int ML>
int launch_logp(const FlowMeta& m, const double* theta, const double* x, long long n, double* logp, cudaStream_t s);
template <int D, int ML>
int launch_local_terms(const FlowMeta& m, const double* theta, const double* x, long long n, const EqParams& e,
                       const double* tang, double* eloc, double* logp, double* grad, double* lap, double* O,
                       long long ldo, cudaStream_t s);
template <int D, int ML>
int launch_transform(const FlowMeta& m, const double* theta, const double* x, long long n, int inv, double* y,
                     double* logjac, double* lat_in, cudaStream_t s);
template <int D, int ML>
int launch_hessian(const FlowMeta& m, const double* theta, const double* x, long long n, double* H, cudaStream_t s);

#define VMC_FOR_EACH_DIM(X) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(12)

}  // namespace vmc
