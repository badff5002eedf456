// HBM-bound passes around the Gram: first moments, centring + force vector, symmetrise/scale, diagonal shift.
// Replaces the reference's tiny pmaps and host-staged reductions: tdvp.py:28-33,37-45,50-51 and
// mpi_wrapper.py:129-245 (local part; the cross-rank sum is one NCCL allreduce of the packed buffer).
// All reductions are single-writer and fixed-order, so results are run-to-run deterministic.
#include <cstdint>
#include "common.cuh"

namespace vmc {

constexpr int kRT = 256;                 // threads per CTA
constexpr int kKC = 8;                   // double2 per thread and row: a CTA covers 2 * kRT * kKC = 4096 columns
constexpr int kTileCols = 2 * kRT * kKC;
constexpr int kRowsPerCta = 512;         // fixed, so that the summation order does not depend on the device

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

// Row-block streaming pass over O (every row is read as one contiguous run: HBM-friendly), two-level fixed-order
// reduction:  MODE 0        :  part[rb][c] = sum_{i in block rb} O[i][c]
//             MODE 1 (CENTER):  O[i][c] -= meanO[c];  part[rb][c] = sum_{i in block rb} (E[i] - meanE) * O[i][c]
// grid = (column tiles of 4096, row blocks of 512); ldo must be even.
//             MODE 2        :  part[rb][c] = sum_{i in block rb} eloc[i] * O[i][c]          (eloc = any per-row factor)
template <int MODE>
__global__ void __launch_bounds__(kRT)
rowblock_kernel(double* __restrict__ O, long long n, long long ldo, const double* __restrict__ meanO,
                const double* __restrict__ eloc, double meanE, double* __restrict__ part) {
  constexpr bool CENTER = MODE == 1;
  const long long c0 = (long long)blockIdx.x * kTileCols + 2 * threadIdx.x;
  const long long i0 = (long long)blockIdx.y * kRowsPerCta;
  const long long i1 = min(n, i0 + kRowsPerCta);
  double2 acc[kKC], mu[kKC];
  bool in[kKC];
#pragma unroll
  for (int k = 0; k < kKC; ++k) {
    const long long c = c0 + (long long)k * 2 * kRT;
    in[k] = c < ldo;
    acc[k] = make_double2(0.0, 0.0);
    mu[k] = (CENTER && in[k]) ? *(const double2*)(meanO + c) : make_double2(0.0, 0.0);
  }
  auto load_row = [&](long long i, double2 (&v)[kKC]) {
    const double* row = O + i * ldo + c0;
#pragma unroll
    for (int k = 0; k < kKC; ++k) v[k] = in[k] ? *(const double2*)(row + k * 2 * kRT) : make_double2(0.0, 0.0);
  };
  auto use_row = [&](long long i, double2 (&v)[kKC]) {
    if (CENTER) {
      double* row = O + i * ldo + c0;
      const double de = __ldg(eloc + i) - meanE;
#pragma unroll
      for (int k = 0; k < kKC; ++k) {
        v[k].x -= mu[k].x; v[k].y -= mu[k].y;
        if (in[k]) *(double2*)(row + k * 2 * kRT) = v[k];
        acc[k].x = fma(de, v[k].x, acc[k].x); acc[k].y = fma(de, v[k].y, acc[k].y);
      }
    } else if (MODE == 2) {
      const double f = __ldg(eloc + i);
#pragma unroll
      for (int k = 0; k < kKC; ++k) { acc[k].x = fma(f, v[k].x, acc[k].x); acc[k].y = fma(f, v[k].y, acc[k].y); }
    } else {
#pragma unroll
      for (int k = 0; k < kKC; ++k) { acc[k].x += v[k].x; acc[k].y += v[k].y; }
    }
  };
  long long i = i0;
  for (; !CENTER && i + 1 < i1; i += 2) {  // column sums: two rows (16 x 16 B per thread) in flight; the centring
                                            // pass keeps one (register pressure costs more than it gains there)
    double2 va[kKC], vb[kKC];
    load_row(i, va);
    load_row(i + 1, vb);
    use_row(i, va);
    use_row(i + 1, vb);
  }
  for (; i < i1; ++i) {
    double2 va[kKC];
    load_row(i, va);
    use_row(i, va);
  }
  double* out = part + (long long)blockIdx.y * ldo + c0;
#pragma unroll
  for (int k = 0; k < kKC; ++k)
    if (in[k]) *(double2*)(out + k * 2 * kRT) = acc[k];
}

// t[i] = w[i] * sum_c O[i][c] v[c]   (warp per row, v through L1; w may be NULL)
__global__ void __launch_bounds__(256) rowdot_kernel(const double* __restrict__ O, long long n, long long ldo, int cols,
                                                     const double* __restrict__ w, const double* __restrict__ v,
                                                     double* __restrict__ t) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long i = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * wpb) {
    const double* row = O + i * ldo;
    double s0 = 0.0, s1 = 0.0;
    int c = 2 * lane;
    for (; c + 64 < cols; c += 128) {
      const double2 a = *(const double2*)(row + c), b = *(const double2*)(row + c + 64);
      const double2 x = __ldg((const double2*)(v + c)), y = __ldg((const double2*)(v + c + 64));
      s0 = fma(a.x, x.x, s0); s0 = fma(a.y, x.y, s0); s1 = fma(b.x, y.x, s1); s1 = fma(b.y, y.y, s1);
    }
    if (c < cols) {
      const double2 a = *(const double2*)(row + c);
      const double2 x = __ldg((const double2*)(v + c));
      s0 = fma(a.x, x.x, s0); s0 = fma(a.y, x.y, s0);
    }
    double sum = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) t[i] = w ? w[i] * sum : sum;
  }
}

// dst[c] += sum_rb part[rb][c]  (fixed order)
__global__ void __launch_bounds__(256) colsum_partials_kernel(const double* __restrict__ part, int row_blocks, long long ldo,
                                                              double* __restrict__ dst) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ldo) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int rb = 0;
  for (; rb + 3 < row_blocks; rb += 4) {
    s0 += part[(long long)rb * ldo + c]; s1 += part[(long long)(rb + 1) * ldo + c];
    s2 += part[(long long)(rb + 2) * ldo + c]; s3 += part[(long long)(rb + 3) * ldo + c];
  }
  for (; rb < row_blocks; ++rb) s0 += part[(long long)rb * ldo + c];
  dst[c] += (s0 + s1) + (s2 + s3);
}

// sums[0..3] += (sum E, sum|E|, sum E^2, sum logp)   (one CTA)
__global__ void __launch_bounds__(1024) scalar_moments_kernel(const double* __restrict__ eloc, const double* __restrict__ logp,
                                                              long long n, double* __restrict__ sums) {
  __shared__ double shs[32];
  double e = 0.0, ea = 0.0, e2 = 0.0, lp = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
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

// dE, dE^2, logp^2 per sample and var_sum += sum dE^2   (one CTA)
__global__ void __launch_bounds__(1024) sample_weights_kernel(const double* __restrict__ eloc, const double* __restrict__ logp,
                                                              long long n, double meanE, double* __restrict__ dE,
                                                              double* __restrict__ wE, double* __restrict__ wLp,
                                                              double* __restrict__ var_sum) {
  __shared__ double shs[32];
  double v2 = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = eloc[i] - meanE;
    if (dE) dE[i] = d;
    if (wE) wE[i] = d * d;
    if (wLp) wLp[i] = logp[i] * logp[i];
    v2 += d * d;
  }
  const double r = block_reduce_sum(v2, shs);
  if (threadIdx.x == 0 && var_sum) var_sum[0] += r;
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

extern "C" __attribute__((visibility("default"))) int vmcpde_moments_workspace_bytes(int64_t n, int64_t ldo, size_t* bytes) {
  using namespace vmc;
  VMC_REQUIRE(bytes && n >= 0 && ldo >= 0, "vmcpde_moments_workspace_bytes: bad arguments");
  const long long row_blocks = (n + kRowsPerCta - 1) / kRowsPerCta;
  *bytes = (size_t)(row_blocks > 0 ? row_blocks : 1) * (size_t)ldo * 8;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_moments1(const double* eloc, const double* logp, const double* O, int64_t n, int64_t ldo,
                               double* sums, void* workspace, size_t workspace_bytes, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(sums, "vmcpde_moments1: null sums");
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (O) {
    VMC_REQUIRE(ldo % 2 == 0 && ((uintptr_t)O & 15) == 0, "vmcpde_moments1: O must be 16-byte aligned with an even ldo");
    size_t need = 0;
    vmcpde_moments_workspace_bytes(n, ldo, &need);
    VMC_REQUIRE(workspace && workspace_bytes >= need, "vmcpde_moments1: workspace too small");
    const int row_blocks = (int)((n + kRowsPerCta - 1) / kRowsPerCta);
    const dim3 grid((unsigned)((ldo + kTileCols - 1) / kTileCols), (unsigned)row_blocks);
    rowblock_kernel<0><<<grid, kRT, 0, s>>>(const_cast<double*>(O), n, ldo, nullptr, nullptr, 0.0, (double*)workspace);
    colsum_partials_kernel<<<(unsigned)((ldo + 255) / 256), 256, 0, s>>>((const double*)workspace, row_blocks, ldo, sums + 4);
  }
  scalar_moments_kernel<<<1, 1024, 0, s>>>(eloc, logp, n, sums);
  VMC_LAUNCH_CHECK("moments1");
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_center_force(double* O, int64_t n, int64_t ldo, const double* meanO, const double* eloc,
                                   const double* logp, double meanE, double* dE, double* wE, double* wLp,
                                   double* Fsum, double* var_sum, void* workspace, size_t workspace_bytes,
                                   vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(O && meanO && eloc && Fsum, "vmcpde_center_force: null pointer");
  VMC_REQUIRE(!wLp || logp, "vmcpde_center_force: wLp requires logp");
  if (n <= 0) return 0;
  VMC_REQUIRE(ldo % 2 == 0 && ((uintptr_t)O & 15) == 0 && ((uintptr_t)meanO & 15) == 0,
              "vmcpde_center_force: O / meanO must be 16-byte aligned with an even ldo");
  size_t need = 0;
  vmcpde_moments_workspace_bytes(n, ldo, &need);
  VMC_REQUIRE(workspace && workspace_bytes >= need, "vmcpde_center_force: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int row_blocks = (int)((n + kRowsPerCta - 1) / kRowsPerCta);
  const dim3 grid((unsigned)((ldo + kTileCols - 1) / kTileCols), (unsigned)row_blocks);
  rowblock_kernel<1><<<grid, kRT, 0, s>>>(O, n, ldo, meanO, eloc, meanE, (double*)workspace);
  colsum_partials_kernel<<<(unsigned)((ldo + 255) / 256), 256, 0, s>>>((const double*)workspace, row_blocks, ldo, Fsum);
  sample_weights_kernel<<<1, 1024, 0, s>>>(eloc, logp, n, meanE, dE, wE, wLp, var_sum);
  VMC_LAUNCH_CHECK("center_force");
  return 0;
}

// Matrix-free product with a weighted Gram: out[c] += sum_i w[i] (O[i,:] . v) O[i,c], i.e. out += (O^T diag(w) O) v without
// forming the P x P matrix (two streaming passes over O instead of N P^2 flops).  Used for quadratic forms with SExp
// (stepper.py:71 reads SExp only through normFunction(v, SExp)).  v, out: ldo doubles (v zero in the padding columns);
// t: n doubles of scratch; workspace as vmcpde_moments1.
extern "C" __attribute__((visibility("default"))) int vmcpde_gram_matvec(const double* O, int64_t n, int64_t ldo, const double* w, const double* v,
                                                                        double* t, double* out, void* workspace,
                                                                        size_t workspace_bytes, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(O && v && t && out, "vmcpde_gram_matvec: null pointer");
  if (n <= 0) return 0;
  VMC_REQUIRE(ldo % 2 == 0 && ((uintptr_t)O & 15) == 0 && ((uintptr_t)v & 15) == 0, "vmcpde_gram_matvec: O / v must be 16-byte aligned with an even ldo");
  size_t need = 0;
  vmcpde_moments_workspace_bytes(n, ldo, &need);
  VMC_REQUIRE(workspace && workspace_bytes >= need, "vmcpde_gram_matvec: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const long long want_blocks = (long long)((n + 7) / 8);
  const int blocks = (int)(want_blocks < (long long)num_sms() * 8 ? want_blocks : (long long)num_sms() * 8);
  rowdot_kernel<<<blocks, 256, 0, s>>>(O, n, ldo, (int)ldo, w, v, t);
  const int row_blocks = (int)((n + kRowsPerCta - 1) / kRowsPerCta);
  const dim3 grid((unsigned)((ldo + kTileCols - 1) / kTileCols), (unsigned)row_blocks);
  rowblock_kernel<2><<<grid, kRT, 0, s>>>(const_cast<double*>(O), n, ldo, nullptr, t, 0.0, (double*)workspace);
  colsum_partials_kernel<<<(unsigned)((ldo + 255) / 256), 256, 0, s>>>((const double*)workspace, row_blocks, ldo, out);
  VMC_LAUNCH_CHECK("gram_matvec");
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
