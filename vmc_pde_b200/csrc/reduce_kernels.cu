// HBM-bound passes around the Gram: first moments, centring + force vector, symmetrise/scale, diagonal shift.
// Replaces the reference's tiny pmaps and host-staged reductions: tdvp.py:28-33,37-45,50-51 and
// mpi_wrapper.py:129-245 (local part; the cross-rank sum is one NCCL allreduce of the packed buffer).
// All reductions are single-writer and fixed-order, so results are run-to-run deterministic.
#include "common.cuh"

namespace vmc {

constexpr int kColThreads = 256;   // 8 warps; lane <-> column, warps stride over rows
constexpr int kColsPerCta = 32;

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
  }
  return r;  // valid on thread 0
}

// sums[4 + c] += sum_i O[i][c] ; last CTA: sums[0..3] += (sum E, sum|E|, sum E^2, sum logp)
__global__ void __launch_bounds__(kColThreads)
moments1_kernel(const double* __restrict__ eloc, const double* __restrict__ logp, const double* __restrict__ O,
                long long n, long long ldo, double* __restrict__ sums, int col_ctas) {
  __shared__ double sh[8][kColsPerCta + 1];
  __shared__ double shs[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x < col_ctas) {
    const long long c = (long long)blockIdx.x * kColsPerCta + lane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (c < ldo) {
      long long i = warp;
      for (; i + 24 < n; i += 32) {
        a0 += O[i * ldo + c];
        a1 += O[(i + 8) * ldo + c];
        a2 += O[(i + 16) * ldo + c];
        a3 += O[(i + 24) * ldo + c];
      }
      for (; i < n; i += 8) a0 += O[i * ldo + c];
    }
    sh[warp][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (warp == 0 && c < ldo) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sh[w][lane];
      sums[4 + c] += s;
    }
  } else {
    double e = 0.0, ea = 0.0, e2 = 0.0, lp = 0.0;
    for (long long i = threadIdx.x; i < n; i += kColThreads) {
      const double v = eloc ? eloc[i] : 0.0;
      e += v; ea += fabs(v); e2 += v * v;
      lp += logp ? logp[i] : 0.0;
    }
    double r;
    r = block_reduce_sum(e, shs);  if (threadIdx.x == 0) sums[0] += r;
    r = block_reduce_sum(ea, shs); if (threadIdx.x == 0) sums[1] += r;
    r = block_reduce_sum(e2, shs); if (threadIdx.x == 0) sums[2] += r;
    r = block_reduce_sum(lp, shs); if (threadIdx.x == 0) sums[3] += r;
  }
}

// O[i][c] -= meanO[c];  Fsum[c] += sum_i (E[i]-meanE) * O[i][c];  last CTA writes dE, wE, wLp, var_sum
__global__ void __launch_bounds__(kColThreads)
center_force_kernel(double* __restrict__ O, long long n, long long ldo, const double* __restrict__ meanO,
                    const double* __restrict__ eloc, const double* __restrict__ logp, double meanE,
                    double* __restrict__ dE, double* __restrict__ wE, double* __restrict__ wLp,
                    double* __restrict__ Fsum, double* __restrict__ var_sum, int col_ctas) {
  __shared__ double sh[8][kColsPerCta + 1];
  __shared__ double shs[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x < col_ctas) {
    const long long c = (long long)blockIdx.x * kColsPerCta + lane;
    double a0 = 0.0, a1 = 0.0;
    if (c < ldo) {
      const double mu = meanO[c];
      long long i = warp;
      for (; i + 8 < n; i += 16) {
        const double v0 = O[i * ldo + c] - mu, v1 = O[(i + 8) * ldo + c] - mu;
        O[i * ldo + c] = v0;
        O[(i + 8) * ldo + c] = v1;
        a0 = fma(eloc[i] - meanE, v0, a0);
        a1 = fma(eloc[i + 8] - meanE, v1, a1);
      }
      for (; i < n; i += 8) {
        const double v0 = O[i * ldo + c] - mu;
        O[i * ldo + c] = v0;
        a0 = fma(eloc[i] - meanE, v0, a0);
      }
    }
    sh[warp][lane] = a0 + a1;
    __syncthreads();
    if (warp == 0 && c < ldo) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sh[w][lane];
      Fsum[c] += s;
    }
  } else {
    double v2 = 0.0;
    for (long long i = threadIdx.x; i < n; i += kColThreads) {
      const double d = eloc[i] - meanE;
      if (dE) dE[i] = d;
      if (wE) wE[i] = d * d;
      if (wLp) wLp[i] = logp[i] * logp[i];
      v2 += d * d;
    }
    const double r = block_reduce_sum(v2, shs);
    if (threadIdx.x == 0 && var_sum) var_sum[0] += r;
  }
}

// upper triangle scaled and mirrored into the lower one, 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) sym_finalize_kernel(double* __restrict__ S, int Pp, double scale) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long idx = (long long)(bi * 32 + r) * Pp + bj * 32 + tx;
    const double v = S[idx] * scale;
    tile[r][tx] = v;
    if (bi != bj || r <= tx) S[idx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // element (bj*32 + r, bi*32 + tx) = transposed (bi*32 + tx, bj*32 + r)
    if (bi != bj || tx < r) S[(long long)(bj * 32 + r) * Pp + bi * 32 + tx] = tile[tx][r];
  }
}

__global__ void diag_shift_kernel(const double* __restrict__ S, double* __restrict__ out, int Pp, int P, double shift,
                                  int copy) {
  const long long total = (long long)Pp * Pp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / Pp), c = (int)(i % Pp);
    double v = S[i];
    if (r == c && r < P) v += shift * v;
    if (copy || r == c) out[i] = v;
  }
}

}  // namespace vmc

extern "C" __attribute__((visibility("default"))) int vmcpde_moments1(const double* eloc, const double* logp, const double* O, int64_t n, int64_t ldo,
                               double* sums, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(sums, "vmcpde_moments1: null sums");
  if (n <= 0) return 0;
  const int col_ctas = O ? (int)((ldo + kColsPerCta - 1) / kColsPerCta) : 0;
  moments1_kernel<<<col_ctas + 1, kColThreads, 0, (cudaStream_t)stream>>>(eloc, logp, O, n, ldo, sums, col_ctas);
  VMC_LAUNCH_CHECK("moments1_kernel");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_center_force(double* O, int64_t n, int64_t ldo, const double* meanO, const double* eloc,
                                   const double* logp, double meanE, double* dE, double* wE, double* wLp,
                                   double* Fsum, double* var_sum, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(O && meanO && eloc && Fsum, "vmcpde_center_force: null pointer");
  VMC_REQUIRE(!wLp || logp, "vmcpde_center_force: wLp requires logp");
  if (n <= 0) return 0;
  const int col_ctas = (int)((ldo + kColsPerCta - 1) / kColsPerCta);
  center_force_kernel<<<col_ctas + 1, kColThreads, 0, (cudaStream_t)stream>>>(O, n, ldo, meanO, eloc, logp, meanE, dE, wE,
                                                                              wLp, Fsum, var_sum, col_ctas);
  VMC_LAUNCH_CHECK("center_force_kernel");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_sym_finalize(double* S, int32_t Pp, double scale, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(S && Pp > 0 && Pp % 32 == 0, "vmcpde_sym_finalize: Pp must be a positive multiple of 32");
  dim3 grid(Pp / 32, Pp / 32);
  sym_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S, Pp, scale);
  VMC_LAUNCH_CHECK("sym_finalize_kernel");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_diag_shift(const double* S, double* S_shifted, int32_t Pp, int32_t P, double shift,
                                 vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(S && S_shifted && Pp > 0 && P <= Pp, "vmcpde_diag_shift: bad arguments");
  diag_shift_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(S, S_shifted, Pp, P, shift, S != S_shifted);
  VMC_LAUNCH_CHECK("diag_shift_kernel");
  return 0;
}
