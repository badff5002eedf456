// Flow kernels for one compile-time dimension (-DVMC_DIM=d): sampler draw, log p, fused local terms
// (forward jets + reverse sweep emitting O rows), Hessian.  One thread per sample; parameters staged in
// shared memory (warp-uniform broadcast reads); O rows leave through a per-warp 32x32 shared-memory
// transpose so that global stores are 256-byte coalesced row segments.
//
// Replaces: sampler.py:25-34,72-86 + var_state.py:76-79 (sample); var_state.py:29,38-43 (eval);
// var_state.py:31-32,55-67 + evolutionEq.py:84-119 (local terms; jax value_and_grad / jacrev(jacfwd)).
#include <cstdlib>
#include "flow_kernels.cuh"
#include "rng.cuh"

#ifndef VMC_DIM
#error "compile with -DVMC_DIM=<dimension>"
#endif

namespace vmc {

constexpr int kThreads = 128;
#ifndef VMC_LT_MINBLOCKS
#define VMC_LT_MINBLOCKS 3   // 168 registers: 12 resident warps per SM instead of 8 (ncu: latency bound at low occupancy)
#endif
constexpr int kStageStride = 33;
constexpr int kStagePerWarp = 32 * kStageStride;
constexpr size_t kMaxThetaSmem = 160 * 1024;

__device__ __forceinline__ const double* stage_theta(const FlowMeta& m, const double* __restrict__ theta,
                                                     double* sth, int use_smem) {
  if (!use_smem) return theta;
  for (int i = threadIdx.x; i < m.P; i += blockDim.x) sth[i] = theta[i];
  __syncthreads();
  return sth;
}

// Emits one O row per lane, in flat order, through a per-warp transpose buffer.
struct WarpEmit {
  double* stage;  // [32][33] doubles of this warp
  double* obase;  // &O[first row of this warp][0]
  long long ldo;
  int nrows, lane, base, cnt, lo;
  __device__ __forceinline__ void flush_window() {
    if (cnt > lo) {
      __syncwarp();
      if (lane >= lo && lane < cnt) {
        for (int r = 0; r < nrows; ++r) obase[r * ldo + base + lane] = stage[r * kStageStride + lane];
      }
      __syncwarp();
    }
  }
  __device__ __forceinline__ void seek(int p) {
    flush_window();
    base = p & ~31; cnt = p & 31; lo = cnt;
  }
  __device__ __forceinline__ void put(double v) {
    stage[lane * kStageStride + cnt] = v;
    if (++cnt == 32) { flush_window(); base += 32; cnt = 0; lo = 0; }
  }
  __device__ __forceinline__ void finish() { flush_window(); lo = cnt; }
};

template <int D, int ML>
__global__ void __launch_bounds__(kThreads)
sample_kernel(const __grid_constant__ FlowMeta m, const double* __restrict__ theta, uint32_t k0, uint32_t k1,
              long long first, long long n, long long n_total, const double* __restrict__ chi2,
              double* __restrict__ x, double* __restrict__ logp, double* __restrict__ zout, int use_smem) {
  extern __shared__ double sth[];
  const double* th = stage_theta(m, theta, sth, use_smem);
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i >= n) return;
  // chol(S), S = L L^T (util.py:21-26; jax multivariate_normal method='cholesky')
  double L[D][D], C[D][D];
  build_L<D>(m, th, L);
#pragma unroll
  for (int j = 0; j < D; ++j) {
#pragma unroll
    for (int a = j; a < D; ++a) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) s += L[a][c] * L[j][c];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= C[a][k] * C[j][k];
      C[a][j] = (a == j) ? sqrt(s) : s / C[j][j];
    }
  }
  double xi[D], z[D], xo[D];
  const unsigned long long total = (unsigned long long)n_total * D;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const unsigned long long e = (unsigned long long)(first + i) * D + j;
    xi[j] = normal_from_bits(random_bits64(k0, k1, e, total));
  }
  double scale = 1.0;
  if (m.latent == kStudentT) {  // sampler.py:29-34
    const double nu = exp(th[m.off_dist]) + 1.0;
    scale = sqrt(nu / chi2[i]);
  }
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k <= a; ++k) s += C[a][k] * xi[k];
    z[a] = (th[m.off_mu + a] + scale * s) + m.offset[a];
  }
  const double lp = sample_from_latent<D>(m, th, z, xo);
#pragma unroll
  for (int a = 0; a < D; ++a) {
    x[i * D + a] = xo[a];
    if (zout) zout[i * D + a] = z[a];
  }
  logp[i] = lp;
}

template <int D, int ML>
__global__ void __launch_bounds__(kThreads)
logp_kernel(const __grid_constant__ FlowMeta m, const double* __restrict__ theta, const double* __restrict__ x,
            long long n, double* __restrict__ logp, int use_smem) {
  extern __shared__ double sth[];
  const double* th = stage_theta(m, theta, sth, use_smem);
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i >= n) return;
  double xi[D];
#pragma unroll
  for (int a = 0; a < D; ++a) xi[a] = x[i * D + a];
  logp[i] = logp_value<D>(m, th, xi);
}

template <int D, int ML>
__global__ void __launch_bounds__(kThreads, VMC_LT_MINBLOCKS)
local_terms_kernel(const __grid_constant__ FlowMeta m, const __grid_constant__ EqParams e,
                   const double* __restrict__ theta, const double* __restrict__ x, long long n,
                   const double* __restrict__ tang, double* __restrict__ eloc, double* __restrict__ logp,
                   double* __restrict__ grad, double* __restrict__ lap, double* __restrict__ O, long long ldo,
                   int use_smem, int theta_smem_doubles) {
  extern __shared__ double sth[];
  const double* th = stage_theta(m, theta, sth, use_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  const long long ic = i < n ? i : n - 1;  // tail lanes shadow the last sample; their stores are masked
  double xi[D], w[D], tg[D * D];
#pragma unroll
  for (int a = 0; a < D; ++a) xi[a] = x[ic * D + a];
  equation_weights<D>(e, w);
  const bool aniso = (e.mode == kDiffusionAniso);
  if (aniso) {
#pragma unroll
    for (int a = 0; a < D * D; ++a) tg[a] = tang[a];
  }
  JetResult<D> r;
  logp_jet<D>(m, th, xi, aniso ? tg : nullptr, w, r);
  if (i < n) {
    if (eloc) eloc[i] = local_term<D>(e, xi, r);
    if (logp) logp[i] = r.logp;
    if (lap) lap[i] = r.lap;
    if (grad) {
#pragma unroll
      for (int a = 0; a < D; ++a) grad[i * D + a] = r.dir[a];
    }
  }
  if (O) {
    const long long row0 = blockIdx.x * (long long)kThreads + warp * 32;
    long long nr = n - row0;
    WarpEmit em;
    em.stage = sth + theta_smem_doubles + warp * kStagePerWarp;
    em.obase = O + (row0 < n ? row0 : 0) * ldo;
    em.ldo = ldo;
    em.nrows = nr <= 0 ? 0 : (nr > 32 ? 32 : (int)nr);
    em.lane = lane; em.base = 0; em.cnt = 0; em.lo = 0;
    logp_reverse<D>(m, th, r.zfin, em, (double*)nullptr);
    // zero the padding columns [P, ldo)
    for (int rr = 0; rr < em.nrows; ++rr)
      for (long long c = m.P + lane; c < ldo; c += 32) em.obase[rr * ldo + c] = 0.0;
  }
}

template <int D, int ML>
__global__ void __launch_bounds__(kThreads)
hessian_kernel(const __grid_constant__ FlowMeta m, const double* __restrict__ theta, const double* __restrict__ x,
               long long n, double* __restrict__ H, int use_smem) {
  extern __shared__ double sth[];
  const double* th = stage_theta(m, theta, sth, use_smem);
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i >= n) return;
  double xi[D], Hl[D * D];
#pragma unroll
  for (int a = 0; a < D; ++a) xi[a] = x[i * D + a];
  logp_hessian<D>(m, th, xi, Hl);
  for (int a = 0; a < D * D; ++a) H[i * D * D + a] = Hl[a];
}

// INN map alone (net.py:168-182): y = INN(x) or INN^-1(x), its log-Jacobian, and log p_lat(x - offset)
template <int D, int ML>
__global__ void __launch_bounds__(kThreads)
transform_kernel(const __grid_constant__ FlowMeta m, const double* __restrict__ theta, const double* __restrict__ x,
                 long long n, int inv, double* __restrict__ y, double* __restrict__ logjac, double* __restrict__ lat_in,
                 int use_smem) {
  extern __shared__ double sth[];
  const double* th = stage_theta(m, theta, sth, use_smem);
  const long long i = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (i >= n) return;
  double z[D], yin[D], lj = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a) { z[a] = x[i * D + a]; yin[a] = z[a] - m.offset[a]; }
  if (lat_in) lat_in[i] = latent_logpdf<D>(m, th, yin);
  if (inv) { for (int b = m.depth - 1; b >= 0; --b) lj += block_inverse_value<D>(m, th, b, z); }
  else { for (int b = 0; b < m.depth; ++b) lj += block_forward_value<D>(m, th, b, z); }
#pragma unroll
  for (int a = 0; a < D; ++a) y[i * D + a] = z[a];
  if (logjac) logjac[i] = lj;
}

// ------------------------------------------------------------------------------------------------
#ifndef VMC_ML
#error "compile with -DVMC_DIM=<d> -DVMC_ML=<0|1>"
#endif
template <class K>
static int prep_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) VMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
static inline unsigned grid_for(long long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

template <int D, int ML>
int launch_sample(const FlowMeta& m, const double* theta, uint32_t k0, uint32_t k1, long long first, long long n,
                  long long n_total, const double* chi2, double* x, double* logp, double* zout, cudaStream_t s) {
  if (n <= 0) return 0;
  const size_t tb = (size_t)m.P * 8;
  const int use = tb <= kMaxThetaSmem;
  const size_t smem = use ? tb : 0;
  if (int rc = prep_smem(sample_kernel<D, ML>, smem)) return rc;
  sample_kernel<D, ML><<<grid_for(n), kThreads, smem, s>>>(m, theta, k0, k1, first, n, n_total, chi2, x, logp, zout, use);
  VMC_LAUNCH_CHECK("sample_kernel");
  return 0;
}
template <int D, int ML>
int launch_logp(const FlowMeta& m, const double* theta, const double* x, long long n, double* logp, cudaStream_t s) {
  if (n <= 0) return 0;
  const size_t tb = (size_t)m.P * 8;
  const int use = tb <= kMaxThetaSmem;
  const size_t smem = use ? tb : 0;
  if (int rc = prep_smem(logp_kernel<D, ML>, smem)) return rc;
  logp_kernel<D, ML><<<grid_for(n), kThreads, smem, s>>>(m, theta, x, n, logp, use);
  VMC_LAUNCH_CHECK("logp_kernel");
  return 0;
}
template <int D, int ML>
int launch_local_terms(const FlowMeta& m, const double* theta, const double* x, long long n, const EqParams& e,
                       const double* tang, double* eloc, double* logp, double* grad, double* lap, double* O,
                       long long ldo, cudaStream_t s) {
  if (n <= 0) return 0;
  const size_t tb = (size_t)m.P * 8;
  // Three CTAs per SM (168 registers) need 3 x (theta + 33 KB of emit staging) of shared memory: larger parameter
  // vectors are read through L1 instead (warp-uniform loads; measured 8 % faster at P = 8187 than 2 CTAs with theta staged).
  static const int theta_smem_env = getenv("VMCPDE_THETA_SMEM") ? atoi(getenv("VMCPDE_THETA_SMEM")) : -1;
  const int use = theta_smem_env >= 0 ? (theta_smem_env && tb <= kMaxThetaSmem) : (tb <= 24 * 1024);
  const int theta_doubles = use ? m.P : 0;
  const size_t smem = (size_t)theta_doubles * 8 + (O ? (kThreads / 32) * kStagePerWarp * 8 : 0);
  if (int rc = prep_smem(local_terms_kernel<D, ML>, smem)) return rc;
  local_terms_kernel<D, ML><<<grid_for(n), kThreads, smem, s>>>(m, e, theta, x, n, tang, eloc, logp, grad, lap, O, ldo,
                                                           use, theta_doubles);
  VMC_LAUNCH_CHECK("local_terms_kernel");
  return 0;
}
template <int D, int ML>
int launch_hessian(const FlowMeta& m, const double* theta, const double* x, long long n, double* H, cudaStream_t s) {
  if (n <= 0) return 0;
  const size_t tb = (size_t)m.P * 8;
  const int use = tb <= kMaxThetaSmem;
  const size_t smem = use ? tb : 0;
  if (int rc = prep_smem(hessian_kernel<D, ML>, smem)) return rc;
  hessian_kernel<D, ML><<<grid_for(n), kThreads, smem, s>>>(m, theta, x, n, H, use);
  VMC_LAUNCH_CHECK("hessian_kernel");
  return 0;
}

template <int D, int ML>
int launch_transform(const FlowMeta& m, const double* theta, const double* x, long long n, int inv, double* y,
                     double* logjac, double* lat_in, cudaStream_t s) {
  if (n <= 0) return 0;
  const size_t tb = (size_t)m.P * 8;
  const int use = tb <= kMaxThetaSmem;
  const size_t smem = use ? tb : 0;
  if (int rc = prep_smem(transform_kernel<D, ML>, smem)) return rc;
  transform_kernel<D, ML><<<grid_for(n), kThreads, smem, s>>>(m, theta, x, n, inv, y, logjac, lat_in, use);
  VMC_LAUNCH_CHECK("transform_kernel");
  return 0;
}
template int launch_transform<VMC_DIM, VMC_ML>(const FlowMeta&, const double*, const double*, long long, int, double*, double*,
                                       double*, cudaStream_t);

template int launch_sample<VMC_DIM, VMC_ML>(const FlowMeta&, const double*, uint32_t, uint32_t, long long, long long, long long,
                                    const double*, double*, double*, double*, cudaStream_t);
template int launch_logp<VMC_DIM, VMC_ML>(const FlowMeta&, const double*, const double*, long long, double*, cudaStream_t);
template int launch_local_terms<VMC_DIM, VMC_ML>(const FlowMeta&, const double*, const double*, long long, const EqParams&,
                                         const double*, double*, double*, double*, double*, double*, long long,
                                         cudaStream_t);
template int launch_hessian<VMC_DIM, VMC_ML>(const FlowMeta&, const double*, const double*, long long, double*, cudaStream_t);

}  // namespace vmc
