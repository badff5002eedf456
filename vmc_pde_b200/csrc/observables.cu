// Observables block of TDVP.__call__ (tdvp.py:143-162): sample moments 1-6, covariance, entropy, max E_loc and
// the ball integrals.  Two-stage fixed-grid reductions (deterministic).  SURVEY section 8(f) rank 1.
#include <cstdint>
#include "common.cuh"
#include "rng.cuh"

namespace vmc {

constexpr int kObsThreads = 256;
constexpr int kObsMaxDim = 16;

__device__ __forceinline__ double obs_wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double obs_wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// stage 1: per-CTA partials of  [sum x_0..x_{d-1}, sum logp, max E]  -> part[cta][d + 2]
__global__ void __launch_bounds__(kObsThreads) obs_first_kernel(const double* __restrict__ x, const double* __restrict__ logp,
                                                                const double* __restrict__ eloc, long long n, int d,
                                                                double* __restrict__ part) {
  __shared__ double sh[8][kObsMaxDim + 2];
  double s[kObsMaxDim + 2];
  for (int a = 0; a < d + 1; ++a) s[a] = 0.0;
  s[d + 1] = -INFINITY;
  for (long long i = blockIdx.x * (long long)kObsThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kObsThreads) {
    for (int a = 0; a < d; ++a) s[a] += x[i * d + a];
    if (logp) s[d] += logp[i];
    if (eloc) s[d + 1] = fmax(s[d + 1], eloc[i]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int a = 0; a < d + 1; ++a) { const double v = obs_wsum(s[a]); if (lane == 0) sh[warp][a] = v; }
  { const double v = obs_wmax(s[d + 1]); if (lane == 0) sh[warp][d + 1] = v; }
  __syncthreads();
  if (threadIdx.x < d + 2) {
    double v = sh[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) v = (threadIdx.x == d + 1) ? fmax(v, sh[w][threadIdx.x]) : v + sh[w][threadIdx.x];
    part[(size_t)blockIdx.x * (d + 2) + threadIdx.x] = v;
  }
}
// final: out[a] = sum over CTAs (max for the last entry)
__global__ void obs_reduce_kernel(const double* __restrict__ part, int ctas, int nvals, int max_index, double* __restrict__ out,
                                  int accumulate) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nvals) return;
  double v = part[a];
  for (int c = 1; c < ctas; ++c) v = (a == max_index) ? fmax(v, part[(size_t)c * nvals + a]) : v + part[(size_t)c * nvals + a];
  if (accumulate) out[a] = (a == max_index) ? fmax(out[a], v) : out[a] + v;
  else out[a] = v;
}

// stage 2: central sums about `mean`:  [cov (d*d), m3 (d), m4 (d), m5 (d), m6 (d)] -> part[cta][d*d + 4d]
__global__ void __launch_bounds__(kObsThreads) obs_central_kernel(const double* __restrict__ x, long long n, int d,
                                                                  const double* __restrict__ mean, double* __restrict__ part) {
  extern __shared__ double shc[];  // 8 * nvals
  const int nvals = d * d + 4 * d;
  double mu[kObsMaxDim], dx[kObsMaxDim];
  double s[kObsMaxDim * kObsMaxDim + 4 * kObsMaxDim];
  for (int a = 0; a < nvals; ++a) s[a] = 0.0;
  for (int a = 0; a < d; ++a) mu[a] = mean[a];
  for (long long i = blockIdx.x * (long long)kObsThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kObsThreads) {
    for (int a = 0; a < d; ++a) dx[a] = x[i * d + a] - mu[a];
    for (int a = 0; a < d; ++a) {
      for (int b = 0; b < d; ++b) s[a * d + b] += dx[a] * dx[b];
      const double d2 = dx[a] * dx[a], d3 = d2 * dx[a];
      s[d * d + a] += d3; s[d * d + d + a] += d2 * d2; s[d * d + 2 * d + a] += d3 * d2; s[d * d + 3 * d + a] += d3 * d3;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int a = 0; a < nvals; ++a) { const double v = obs_wsum(s[a]); if (lane == 0) shc[warp * nvals + a] = v; }
  __syncthreads();
  for (int a = threadIdx.x; a < nvals; a += kObsThreads) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += shc[w * nvals + a];
    part[(size_t)blockIdx.x * nvals + a] = v;
  }
}

// tdvp.py:154-155: s = normal(key,(n,d)); s = s/|s| * uniform(key,(n,))**(1/d) * radius   (both draws from the SAME key)
__global__ void ball_points_kernel(uint32_t k0, uint32_t k1, long long first, long long n, long long n_total, int d,
                                   double radius, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[kObsMaxDim], nrm = 0.0;
  for (int a = 0; a < d; ++a) {
    v[a] = normal_from_bits(random_bits64(k0, k1, (unsigned long long)(first + i) * d + a, (unsigned long long)n_total * d));
    nrm += v[a] * v[a];
  }
  const double u = bits_to_unit(random_bits64(k0, k1, (unsigned long long)(first + i), (unsigned long long)n_total));
  const double f = radius * pow(u, 1.0 / d) / sqrt(nrm);
  for (int a = 0; a < d; ++a) out[i * d + a] = f * v[a];
}

// part[cta] = sum exp(logp)
__global__ void __launch_bounds__(kObsThreads) sum_exp_kernel(const double* __restrict__ logp, long long n, double* __restrict__ part) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)kObsThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kObsThreads) s += exp(logp[i]);
  s = obs_wsum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double v = 0.0; for (int w = 0; w < 8; ++w) v += sh[w]; part[blockIdx.x] = v; }
}

}  // namespace vmc
using namespace vmc;

extern "C" __attribute__((visibility("default"))) int vmcpde_observables_workspace_bytes(int32_t d, size_t* bytes) {
  VMC_REQUIRE(bytes && d >= 1 && d <= kObsMaxDim, "vmcpde_observables_workspace_bytes: bad dimension");
  *bytes = (size_t)num_sms() * 2 * (d * d + 4 * d + 2) * 8;
  return 0;
}
// first[0..d) += sum x ; first[d] += sum logp ; first[d+1] = max(first[d+1], max E)     (local sums, tdvp.py:144,147,150)
extern "C" __attribute__((visibility("default"))) int vmcpde_obs_first(const double* x, const double* logp, const double* eloc, int64_t n,
                                                                      int32_t d, double* first, void* ws, vmcpde_stream stream) {
  VMC_REQUIRE(x && first && ws && d >= 1 && d <= kObsMaxDim, "vmcpde_obs_first: bad arguments");
  if (n <= 0) return 0;
  const int ctas = num_sms() * 2;
  obs_first_kernel<<<ctas, kObsThreads, 0, (cudaStream_t)stream>>>(x, logp, eloc, n, d, (double*)ws);
  obs_reduce_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const double*)ws, ctas, d + 2, d + 1, first, 1);
  VMC_LAUNCH_CHECK("obs_first");
  return 0;
}
// central[0..d*d) += sum dx dx^T ; then d entries each of sum dx^3, dx^4, dx^5, dx^6 about `mean` (tdvp.py:146,148-149)
extern "C" __attribute__((visibility("default"))) int vmcpde_obs_central(const double* x, int64_t n, int32_t d, const double* mean,
                                                                        double* central, void* ws, vmcpde_stream stream) {
  VMC_REQUIRE(x && mean && central && ws && d >= 1 && d <= kObsMaxDim, "vmcpde_obs_central: bad arguments");
  if (n <= 0) return 0;
  const int ctas = num_sms() * 2, nvals = d * d + 4 * d;
  obs_central_kernel<<<ctas, kObsThreads, 8 * nvals * 8, (cudaStream_t)stream>>>(x, n, d, mean, (double*)ws);
  obs_reduce_kernel<<<(nvals + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const double*)ws, ctas, nvals, -1, central, 1);
  VMC_LAUNCH_CHECK("obs_central");
  return 0;
}
extern "C" __attribute__((visibility("default"))) int vmcpde_ball_points(uint32_t key0, uint32_t key1, int64_t first, int64_t n,
                                                                        int64_t n_total, int32_t d, double radius, double* out,
                                                                        vmcpde_stream stream) {
  VMC_REQUIRE(out && d >= 1 && d <= kObsMaxDim && first >= 0 && first + n <= n_total, "vmcpde_ball_points: bad arguments");
  VMC_REQUIRE((unsigned long long)n_total * d * 2ull <= 0x100000000ull, "vmcpde_ball_points: counter range exceeded");
  if (n <= 0) return 0;
  ball_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key0, key1, first, n, n_total, d, radius, out);
  VMC_LAUNCH_CHECK("ball_points_kernel");
  return 0;
}
// out[0] += sum_i exp(logp[i])   (tdvp.py:162)
extern "C" __attribute__((visibility("default"))) int vmcpde_sum_exp(const double* logp, int64_t n, double* out, void* ws, vmcpde_stream stream) {
  VMC_REQUIRE(logp && out && ws, "vmcpde_sum_exp: null pointer");
  if (n <= 0) return 0;
  const int ctas = num_sms() * 2;
  sum_exp_kernel<<<ctas, kObsThreads, 0, (cudaStream_t)stream>>>(logp, n, (double*)ws);
  obs_reduce_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const double*)ws, ctas, 1, -1, out, 1);
  VMC_LAUNCH_CHECK("sum_exp");
  return 0;
}
