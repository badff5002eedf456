// Regularised solve of the TDVP equation on the device.
//  * vmcpde_solve_tail: everything after the eigendecomposition in tdvp.py:66-94 -- V^T F, the signal-to-noise
//    ratio (tdvp.py:68-71, computed as diag(V^T C V) with C the dE^2-weighted Gram instead of the reference's
//    N x P x P product EOdata @ V), eigenvalue cut-offs, update = V (invEv * reg * V^T F), solver residual
//    and TDVP error.
//  * vmcpde_chol_solve: blocked Cholesky fast path for a shifted (positive definite) S
//    (north-star item 4; admissible only with diagonalShift > 0).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

extern "C" int vmcpde_gemm_tn(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Out, int64_t ldo,
                              int32_t M, int32_t N, int64_t K, double alpha, double beta, vmcpde_stream stream);
extern "C" int vmcpde_syrk_tn(const double* X, int64_t ldx, double* Out, int64_t ldo, int32_t M, int64_t K, double alpha,
                              double beta, vmcpde_stream stream);

namespace vmc {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double bsum(double v, double* sh) {  // broadcast block sum, sh >= 33 doubles
  v = wsum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (lane < nw) ? sh[lane] : 0.0;
  return wsum(r);
}

// y[r] = sum_c A[r][c] x[c]   (warp per row)
__global__ void __launch_bounds__(256) gemv_rows_kernel(const double* __restrict__ A, int ld, int m, int n,
                                                        const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < m; r += gridDim.x * wpb) {
    const double* row = A + (size_t)r * ld;
    double s0 = 0.0, s1 = 0.0;
    int c = lane;
    for (; c + 32 < n; c += 64) { s0 = fma(row[c], x[c], s0); s1 = fma(row[c + 32], x[c + 32], s1); }
    if (c < n) s0 = fma(row[c], x[c], s0);
    const double s = wsum(s0 + s1);
    if (lane == 0) y[r] = s;
  }
}

// y[c] = sum_r A[r][c] x[r]   (CTA owns 32 columns, 8 warps stride over rows)
__global__ void __launch_bounds__(256) gemv_cols_kernel(const double* __restrict__ A, int ld, int m, int n,
                                                        const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double sh[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  double a0 = 0.0, a1 = 0.0;
  if (c < n) {
    int r = warp;
    for (; r + 8 < m; r += 16) { a0 = fma(A[(size_t)r * ld + c], x[r], a0); a1 = fma(A[(size_t)(r + 8) * ld + c], x[r + 8], a1); }
    if (r < m) a0 = fma(A[(size_t)r * ld + c], x[r], a0);
  }
  sh[warp][lane] = a0 + a1;
  __syncthreads();
  if (warp == 0 && c < n) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][lane];
    y[c] = s;
  }
}

// out[c] = sum_r A[r][c] * B[r][c]
__global__ void __launch_bounds__(256) coldot_kernel(const double* __restrict__ A, const double* __restrict__ B, int ld,
                                                     int m, int n, double* __restrict__ out) {
  __shared__ double sh[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  double a0 = 0.0;
  if (c < n)
    for (int r = warp; r < m; r += 8) a0 = fma(A[(size_t)r * ld + c], B[(size_t)r * ld + c], a0);
  sh[warp][lane] = a0;
  __syncthreads();
  if (warp == 0 && c < n) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][lane];
    out[c] = s;
  }
}

// B[c][r] = A[r][c] on the ld x ld padded square (32x32 tiles)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ A, double* __restrict__ B, int ld) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = A[(size_t)(r0 + r) * ld + c0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) B[(size_t)(c0 + r) * ld + r0 + tx] = tile[tx][r];
}

// rows [r0, r0 + 32 * gridDim.y) of A become the same columns of B (B[c][r] = A[r][c], c < ld)
__global__ void __launch_bounds__(256) transpose_rows_kernel(const double* __restrict__ A, double* __restrict__ B, int ld, int r0) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rr = r0 + blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = A[(size_t)(rr + r) * ld + c0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) B[(size_t)(c0 + r) * ld + rr + tx] = tile[tx][r];
}

// tdvp.py:70-71,82-89: rhoVar, snr, invEv, regulariser -> coefficient in the eigenbasis
__global__ void solve_coef_kernel(const double* __restrict__ ev, const double* __restrict__ VtF,
                                  const double* __restrict__ q, int n, double n_glob, double svdTol, double snrTol,
                                  int useSNR, double* __restrict__ rhoVar, double* __restrict__ snr,
                                  double* __restrict__ invEv, double* __restrict__ coef, int k0, int k1) {
  const double evmax = ev[n - 1];
  for (int k = k0 + blockIdx.x * blockDim.x + threadIdx.x; k < k1; k += gridDim.x * blockDim.x) {
    const double f = VtF[k];
    double sn = 0.0;
    if (q) {
      const double rv = q[k] - f * f;  // population variance of (dE * dO) . v_k
      rhoVar[k] = rv;
      sn = sqrt(fabs(n_glob * f * f / rv));
      snr[k] = sn;
    }
    const double r = fabs(ev[k] / evmax);
    const double inv = r > 1e-14 ? 1.0 / ev[k] : 0.0;
    const double t = svdTol / r, t2 = t * t;
    double reg = 1.0 / (1.0 + t2 * t2 * t2);
    if (useSNR) { const double u = snrTol / sn, u2 = u * u; reg *= 1.0 / (1.0 + u2 * u2 * u2); }
    invEv[k] = inv;
    coef[k] = inv * reg * f;
  }
}

// scalars[0] = ||S u - F|| / ||F|| ; scalars[1] = 1 + (u.S0u - 2 F0.u) / meanE2   (tdvp.py:93-94)
__global__ void __launch_bounds__(1024) solve_scalars_kernel(const double* __restrict__ Su, const double* __restrict__ S0u,
                                                             const double* __restrict__ F, const double* __restrict__ u,
                                                             int n, double meanE2, double* __restrict__ scalars) {
  __shared__ double sh[33];
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = Su[i] - F[i];
    a += r * r; b += F[i] * F[i]; c += u[i] * S0u[i]; d += F[i] * u[i];
  }
  a = bsum(a, sh); b = bsum(b, sh); c = bsum(c, sh); d = bsum(d, sh);
  if (threadIdx.x == 0) {
    scalars[0] = sqrt(a) / sqrt(b);
    scalars[1] = 1.0 + (c - 2.0 * d) / meanE2;
  }
}

// A parameter no sample depends on gives an exactly zero row and column of the Gram, and the multiplicative shift of
// tdvp.py:50-51 adds shift * 0 to its diagonal.  Such a diagonal entry is set to 1 before factorising, so the parameter
// gets the update F_i / 1 = 0 -- what the eigen-solve assigns to a null direction -- instead of failing the factorisation.
__global__ void zero_diag_to_one_kernel(double* __restrict__ A, int ld, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && A[(size_t)i * ld + i] == 0.0) A[(size_t)i * ld + i] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// Blocked Cholesky (lower), nb = 64
constexpr int kNb = 64;

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int ld, int j0, int nb, int* info) {
  __shared__ double L[kNb][kNb + 1];
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    const int r = idx / nb, c = idx % nb;
    L[r][c] = A[(size_t)(j0 + r) * ld + j0 + c];
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (threadIdx.x == 0) {
      const double p = L[j][j];
      if (!(p > 0.0)) { if (*info == 0) *info = j0 + j + 1; L[j][j] = 1.0; } else L[j][j] = sqrt(p);
    }
    __syncthreads();
    const double d = L[j][j];
    for (int i = j + 1 + threadIdx.x; i < nb; i += blockDim.x) L[i][j] /= d;
    __syncthreads();
    const int m = nb - j - 1;
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
      const int i = j + 1 + idx / m, k = j + 1 + idx % m;
      if (k <= i) L[i][k] -= L[i][j] * L[k][j];
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    const int r = idx / nb, c = idx % nb;
    A[(size_t)(j0 + r) * ld + j0 + c] = (c <= r) ? L[r][c] : 0.0;
  }
}

// rows below the diagonal block: X L11^T = A21, thread per row
__global__ void __launch_bounds__(64) trsm_panel_kernel(double* __restrict__ A, int ld, int n, int j0, int nb) {
  __shared__ double L[kNb][kNb + 1];
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) L[idx / nb][idx % nb] = A[(size_t)(j0 + idx / nb) * ld + j0 + idx % nb];
  __syncthreads();
  const int r = j0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double x[kNb];
  double* row = A + (size_t)r * ld + j0;
#pragma unroll 4
  for (int c = 0; c < nb; ++c) {
    double s = row[c];
    for (int k = 0; k < c; ++k) s -= x[k] * L[c][k];
    x[c] = s / L[c][c];
  }
  for (int c = 0; c < nb; ++c) row[c] = x[c];
}

// trailing update, lower tiles: A22[r][c] -= sum_k L21[r][k] L21[c][k]   (64x64 tiles, DMMA, K = nb)
__device__ __forceinline__ void dmma_s(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) syrk_update_kernel(double* __restrict__ A, int ld, int n, int j0, int nb) {
  constexpr int KC = 32;
  __shared__ double Ls[64][KC + 4], Rs[64][KC + 4];
  const int t0 = j0 + nb;
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tc > tr) return;
  const int r0 = t0 + tr * 64, c0 = t0 + tc * 64;
  if (r0 >= n) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
  for (int kk = 0; kk < nb; kk += KC) {
    for (int idx = threadIdx.x; idx < 64 * KC; idx += blockDim.x) {
      const int r = idx / KC, k = idx % KC;
      const bool kin = kk + k < nb;
      Ls[r][k] = (kin && r0 + r < n) ? A[(size_t)(r0 + r) * ld + j0 + kk + k] : 0.0;
      Rs[r][k] = (kin && c0 + r < n) ? A[(size_t)(c0 + r) * ld + j0 + kk + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < KC; k0 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = Ls[wm * 32 + q * 8 + g][k0 + t]; b[q] = Rs[wn * 32 + q * 8 + g][k0 + t]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p) dmma_s(acc[q][p][0], acc[q][p][1], a[q], b[p]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = r0 + wm * 32 + q * 8 + g;
    if (r >= n) continue;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int c = c0 + wn * 32 + p * 8 + 2 * t;
      if (c < n && c <= r) A[(size_t)r * ld + c] -= acc[q][p][0];
      if (c + 1 < n && c + 1 <= r) A[(size_t)r * ld + c + 1] -= acc[q][p][1];
    }
  }
}

// L y = b then L^T x = y, single CTA; diagonal blocks are staged in shared memory and solved by warp 0
__global__ void __launch_bounds__(1024) chol_substitute_kernel(const double* __restrict__ L, int ld, int n,
                                                               const double* __restrict__ b, double* __restrict__ x) {
  __shared__ double blk[kNb];
  __shared__ double Ld[kNb][kNb + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = b[i];
  __syncthreads();
  for (int j0 = 0; j0 < n; j0 += kNb) {  // forward
    const int nb = min(kNb, n - j0);
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) Ld[idx / nb][idx % nb] = L[(size_t)(j0 + idx / nb) * ld + j0 + idx % nb];
    if ((int)threadIdx.x < nb) blk[threadIdx.x] = x[j0 + threadIdx.x];
    __syncthreads();
    if (warp == 0) {
      for (int r = 0; r < nb; ++r) {
        double s = 0.0;
        for (int k = lane; k < r; k += 32) s += Ld[r][k] * blk[k];
        s = wsum(s);
        if (lane == 0) blk[r] = (blk[r] - s) / Ld[r][r];
        __syncwarp();
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) x[j0 + threadIdx.x] = blk[threadIdx.x];
    for (int r = j0 + nb + threadIdx.x; r < n; r += blockDim.x) {
      const double* row = L + (size_t)r * ld + j0;
      double s = 0.0;
      for (int k = 0; k < nb; ++k) s = fma(row[k], blk[k], s);
      x[r] -= s;
    }
    __syncthreads();
  }
  for (int j0 = ((n - 1) / kNb) * kNb; j0 >= 0; j0 -= kNb) {  // backward with L^T
    const int nb = min(kNb, n - j0);
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) Ld[idx / nb][idx % nb] = L[(size_t)(j0 + idx / nb) * ld + j0 + idx % nb];
    if ((int)threadIdx.x < nb) blk[threadIdx.x] = x[j0 + threadIdx.x];
    __syncthreads();
    if (warp == 0) {
      for (int r = nb - 1; r >= 0; --r) {
        double s = 0.0;
        for (int k = r + 1 + lane; k < nb; k += 32) s += Ld[k][r] * blk[k];
        s = wsum(s);
        if (lane == 0) blk[r] = (blk[r] - s) / Ld[r][r];
        __syncwarp();
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) x[j0 + threadIdx.x] = blk[threadIdx.x];
    for (int c = threadIdx.x; c < j0; c += blockDim.x) {
      double s = 0.0;
      for (int k = 0; k < nb; ++k) s = fma(L[(size_t)(j0 + k) * ld + c], blk[k], s);
      x[c] -= s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core blocked Cholesky (upper form A = U^T U, nb = 128) for padded matrices (ld % 128 == 0).  Per panel:
//   potrf128_kernel   U11 of the 128 x 128 diagonal block (one CTA, shared memory);
//   trsm128_kernel    U12 = U11^-T A12 in place, one thread per column of A12 (coalesced rows, U11 broadcast from
//                     shared memory, 32 right-hand-side entries in registers at a time);
//   vmcpde_syrk_tn    A22 -= U12^T U12 on the upper-triangular tiles (DMMA pipeline of the Gram kernel).
// U lives in the upper triangle; the lower triangle is never read.
constexpr int kCb = 128;

__global__ void __launch_bounds__(256) potrf128_kernel(double* __restrict__ A, int ld, int j0, int n, int* info) {
  // Register-tiled right-looking factorisation: thread (ty, tx) of a 16 x 16 grid owns the entries (ty + 16 i, tx + 16 j)
  // of the block (cyclic, so the shrinking trailing part stays balanced); per step the scaled pivot row goes through
  // shared memory and every thread applies the rank-1 update to its 8 x 8 registers.
  __shared__ double rowk[kCb];
  __shared__ double piv;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double u[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      double v = (c >= r && j0 + r < n && j0 + c < n) ? A[(size_t)(j0 + r) * ld + j0 + c] : 0.0;
      if ((j0 + r >= n || j0 + c >= n) && r == c) v = 1.0;   // identity in the padding
      u[i][j] = v;
    }
  for (int k = 0; k < kCb; ++k) {
    const int ki = k >> 4, kt = k & 15;   // row k lives in register row ki of the threads with ty == kt (same for columns)
    if (ty == kt && tx == kt) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i == ki) piv = u[i][i];
    }
    __syncthreads();
    const double p = piv;
    double d = 1.0;
    if (!(p > 0.0)) { if (tid == 0 && *info == 0) *info = j0 + k + 1; } else d = sqrt(p);
    const double dinv = 1.0 / d;
    if (ty == kt) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i == ki) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = tx + 16 * j;
            if (c > k) u[i][j] *= dinv; else if (c == k) u[i][j] = d;
            rowk[c] = u[i][j];
          }
        }
    }
    __syncthreads();
    double rk[8], ck[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rk[i] = rowk[ty + 16 * i];
#pragma unroll
    for (int j = 0; j < 8; ++j) ck[j] = rowk[tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = ty + 16 * i, c = tx + 16 * j;
        if (r > k && c >= r) u[i][j] = fma(-rk[i], ck[j], u[i][j]);
      }
    // the next step's pivot write happens before its barrier; rowk is rewritten only after that barrier
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      if (j0 + r < n && j0 + c < n && c >= r) A[(size_t)(j0 + r) * ld + j0 + c] = u[i][j];
    }
}

// A12 (128 rows j0.., m columns from j1) <- U11^-T A12: forward substitution down each column
__global__ void __launch_bounds__(128) trsm128_kernel(double* __restrict__ A, int ld, int j0, int m) {
  extern __shared__ double sm[];            // U11 [128][129]
  double* U = sm;
  for (int idx = threadIdx.x; idx < kCb * kCb; idx += 128) {
    const int r = idx >> 7, c = idx & 127;
    U[r * (kCb + 1) + c] = (c >= r) ? A[(size_t)(j0 + r) * ld + j0 + c] : 0.0;
  }
  __syncthreads();
  const int col = blockIdx.x * 128 + threadIdx.x;
  if (col >= m) return;
  double* a = A + (size_t)j0 * ld + j0 + kCb + col;   // a[k * ld] = A12[k][col]
  for (int kc = 0; kc < kCb; kc += 32) {
    double y[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) y[k] = a[(size_t)(kc + k) * ld];
    for (int q = 0; q < kc; ++q) {            // contributions of the finished chunks
      const double yq = a[(size_t)q * ld];
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] = fma(-U[q * (kCb + 1) + kc + k], yq, y[k]);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {            // inside the chunk
      y[k] /= U[(kc + k) * (kCb + 1) + kc + k];
#pragma unroll
      for (int k2 = k + 1; k2 < 32; ++k2) y[k2] = fma(-U[(kc + k) * (kCb + 1) + kc + k2], y[k], y[k2]);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) a[(size_t)(kc + k) * ld] = y[k];
  }
}

// Substitution U^T y = b, U x = y (U upper, rows contiguous), block by block: the 128 x 128 diagonal solves run in one CTA
// out of shared memory, the off-diagonal parts are streamed by the whole GPU.
// forward diagonal block: x[j0..] <- U11^-T x[j0..]
__global__ void __launch_bounds__(128) chol_diag_forward_kernel(const double* __restrict__ U, int ld, int n, int j0, double* __restrict__ x) {
  extern __shared__ double sm[];            // Ud [128][129], blk [128]
  double* Ud = sm;
  double* blk = sm + kCb * (kCb + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbk = min(kCb, n - j0);
  for (int idx = tid; idx < kCb * kCb; idx += 128) {
    const int r = idx >> 7, c = idx & 127;
    Ud[r * (kCb + 1) + c] = (r < nbk && c < nbk && c >= r) ? U[(size_t)(j0 + r) * ld + j0 + c] : (r == c ? 1.0 : 0.0);
  }
  blk[tid] = tid < nbk ? x[j0 + tid] : 0.0;
  __syncthreads();
  if (warp == 0) {
    for (int r = 0; r < nbk; ++r) {
      const double yr = blk[r] / Ud[r * (kCb + 1) + r];
      __syncwarp();
      if (lane == 0) blk[r] = yr;
      for (int c = r + 1 + lane; c < nbk; c += 32) blk[c] = fma(-Ud[r * (kCb + 1) + c], yr, blk[c]);
      __syncwarp();
    }
  }
  __syncthreads();
  if (tid < nbk) x[j0 + tid] = blk[tid];
}
// forward off-block update: x[c] -= sum_r U[j0 + r][c] y[j0 + r] for c >= j0 + 128
__global__ void __launch_bounds__(256) chol_offblock_forward_kernel(const double* __restrict__ U, int ld, int n, int j0, double* __restrict__ x) {
  __shared__ double yb[kCb];
  if (threadIdx.x < kCb) yb[threadIdx.x] = x[j0 + threadIdx.x];
  __syncthreads();
  const int c = j0 + kCb + blockIdx.x * 256 + threadIdx.x;
  if (c >= n) return;
  double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
  for (int r = 0; r < kCb; r += 2) {
    s0 = fma(U[(size_t)(j0 + r) * ld + c], yb[r], s0);
    s1 = fma(U[(size_t)(j0 + r + 1) * ld + c], yb[r + 1], s1);
  }
  x[c] -= s0 + s1;
}
// backward: x[j0 + r] <- (x[j0 + r] - sum_{c >= j0 + 128} U[j0 + r][c] x[c]) then the diagonal block solve.  The row dots
// run as one warp per row over 16 CTAs, the block solve in the last-arriving... (kept simple: two kernels)
__global__ void __launch_bounds__(256) chol_offblock_backward_kernel(const double* __restrict__ U, int ld, int n, int j0, double* __restrict__ x) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int nbk = min(kCb, n - j0);
  if (r >= nbk) return;
  const double* row = U + (size_t)(j0 + r) * ld;
  double s0 = 0.0, s1 = 0.0;
  int c = j0 + kCb + lane;
  for (; c + 32 < n; c += 64) { s0 = fma(row[c], x[c], s0); s1 = fma(row[c + 32], x[c + 32], s1); }
  if (c < n) s0 = fma(row[c], x[c], s0);
  const double sacc = wsum(s0 + s1);
  if (lane == 0) x[j0 + r] -= sacc;
}
__global__ void __launch_bounds__(128) chol_diag_backward_kernel(const double* __restrict__ U, int ld, int n, int j0, double* __restrict__ x) {
  extern __shared__ double sm[];
  double* Ud = sm;
  double* blk = sm + kCb * (kCb + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbk = min(kCb, n - j0);
  for (int idx = tid; idx < kCb * kCb; idx += 128) {
    const int r = idx >> 7, c = idx & 127;
    Ud[r * (kCb + 1) + c] = (r < nbk && c < nbk && c >= r) ? U[(size_t)(j0 + r) * ld + j0 + c] : (r == c ? 1.0 : 0.0);
  }
  blk[tid] = tid < nbk ? x[j0 + tid] : 0.0;
  __syncthreads();
  if (warp == 0) {
    for (int r = nbk - 1; r >= 0; --r) {
      double sacc = 0.0;
      for (int c = r + 1 + lane; c < nbk; c += 32) sacc = fma(Ud[r * (kCb + 1) + c], blk[c], sacc);
      sacc = wsum(sacc);
      if (lane == 0) blk[r] = (blk[r] - sacc) / Ud[r * (kCb + 1) + r];
      __syncwarp();
    }
  }
  __syncthreads();
  if (tid < nbk) x[j0 + tid] = blk[tid];
}

static size_t al(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace vmc

using namespace vmc;

extern "C" __attribute__((visibility("default"))) int vmcpde_solve_tail_workspace_bytes(int32_t n, int32_t ld, size_t* bytes) {
  VMC_REQUIRE(bytes && n >= 1 && ld >= n, "vmcpde_solve_tail_workspace_bytes: bad arguments");
  *bytes = 2 * al((size_t)ld * ld * 8) + 6 * al((size_t)ld * 8);
  return 0;
}

// The eigenvector-local part of TDVP.solve / transform_to_eigenbasis (tdvp.py:66-92) for the eigenvectors
// [row0, row0 + nrows) (rows of VT; multiples of 128 when CEO is given): VtF, rhoVar, snr, invEv on that range and
// update_partial[n] = sum_{k in range} V[:,k] invEv_k reg_k VtF_k.  A multi-GPU solve gives every rank a slice and sums
// update_partial (and the range vectors, zero elsewhere) with one all-reduce; one rank with the full range gets
// the reference's result.  Entries of the output vectors outside the range are not written.
extern "C" __attribute__((visibility("default"))) int vmcpde_solve_tail_range(
    const double* ev, const double* VT, int32_t n, int32_t ld, const double* F, const double* CEO, double n_glob,
    double svdTol, double snrTol, int32_t useSNR, int32_t row0, int32_t nrows, double* VtF, double* rhoVar, double* snr,
    double* invEv, double* update_partial, void* workspace, size_t workspace_bytes, vmcpde_stream stream) {
  VMC_REQUIRE(ev && VT && F && VtF && invEv && update_partial && workspace, "vmcpde_solve_tail_range: null pointer");
  VMC_REQUIRE(!useSNR || CEO, "vmcpde_solve_tail_range: useSNR needs the SNR covariance");
  VMC_REQUIRE(!CEO || (rhoVar && snr), "vmcpde_solve_tail_range: rhoVar/snr outputs required with CEO");
  VMC_REQUIRE(!CEO || (ld % 128 == 0 && row0 % 128 == 0 && nrows % 128 == 0), "vmcpde_solve_tail_range: ld, row0, nrows must be multiples of 128");
  VMC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= ld, "vmcpde_solve_tail_range: bad range");
  size_t need = 0;
  vmcpde_solve_tail_workspace_bytes(n, ld, &need);
  VMC_REQUIRE(workspace_bytes >= need, "vmcpde_solve_tail_range: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* wp = (uint8_t*)workspace;
  double* V = (double*)wp; wp += al((size_t)ld * ld * 8);
  double* W = (double*)wp; wp += al((size_t)ld * ld * 8);
  double* q = (double*)wp; wp += al((size_t)ld * 8);
  double* coef = (double*)wp; wp += al((size_t)ld * 8);
  const int sms = num_sms();
  const int k0 = row0, k1 = min(n, row0 + nrows), nk = k1 - k0;  // real eigenvectors of the range
  if (nk <= 0) {
    VMC_CUDA_CHECK(cudaMemsetAsync(update_partial, 0, (size_t)n * 8, s));
    return 0;
  }
  const int rb = max(1, min(sms * 4, (nk + 7) / 8));
  gemv_rows_kernel<<<rb, 256, 0, s>>>(VT + (size_t)k0 * ld, ld, nk, n, F, VtF + k0);
  if (CEO) {
    transpose_rows_kernel<<<dim3(ld / 32, nrows / 32), 256, 0, s>>>(VT, V, ld, row0);
    if (int rc = vmcpde_gemm_tn(CEO, ld, V + row0, ld, W + row0, ld, ld, nrows, ld, 1.0, 0.0, stream)) return rc;
    coldot_kernel<<<(nk + 31) / 32, 256, 0, s>>>(V + k0, W + k0, ld, n, nk, q + k0);
  }
  solve_coef_kernel<<<max(1, min(sms, (nk + 255) / 256)), 256, 0, s>>>(ev, VtF, CEO ? q : nullptr, n, n_glob, svdTol, snrTol,
                                                                     useSNR, rhoVar, snr, invEv, coef, k0, k1);
  gemv_cols_kernel<<<(n + 31) / 32, 256, 0, s>>>(VT + (size_t)k0 * ld, ld, nk, n, coef + k0, update_partial);
  VMC_LAUNCH_CHECK("solve_tail_range");
  return 0;
}

// Everything after eigh in TDVP.solve / transform_to_eigenbasis (tdvp.py:66-94).
// ev[n] ascending, VT[n x ld] rows = eigenvectors (pad region zero), F[n], S (shifted) and S0 [n x ld] full symmetric,
// CEO [ld x ld] = (1/N) sum dE^2 dO dO^T (padded, zero pad) or NULL to skip the SNR.
// Outputs: VtF, rhoVar, snr (NULL-able when CEO is NULL), invEv, update (n each), scalars[2] = {residual, tdvp_error}.
extern "C" __attribute__((visibility("default"))) int vmcpde_solve_tail(
    const double* ev, const double* VT, int32_t n, int32_t ld, const double* F, const double* S, const double* S0,
    const double* CEO, double n_glob, double svdTol, double snrTol, int32_t useSNR, double meanE2, double* VtF,
    double* rhoVar, double* snr, double* invEv, double* update, double* scalars, void* workspace,
    size_t workspace_bytes, vmcpde_stream stream) {
  VMC_REQUIRE(S && S0 && scalars && update && workspace, "vmcpde_solve_tail: null pointer");
  VMC_REQUIRE(!CEO || ld % 128 == 0, "vmcpde_solve_tail: ld must be a multiple of 128");
  const int nrows = CEO ? ld : n;
  if (int rc = vmcpde_solve_tail_range(ev, VT, n, ld, F, CEO, n_glob, svdTol, snrTol, useSNR, 0, nrows, VtF, rhoVar, snr, invEv,
                                       update, workspace, workspace_bytes, stream))
    return rc;
  uint8_t* wp = (uint8_t*)workspace + 2 * al((size_t)ld * ld * 8) + 2 * al((size_t)ld * 8);
  double* Su = (double*)wp; wp += al((size_t)ld * 8);
  double* S0u = (double*)wp;
  cudaStream_t s = (cudaStream_t)stream;
  const int rb = max(1, min(num_sms() * 4, (n + 7) / 8));
  gemv_rows_kernel<<<rb, 256, 0, s>>>(S, ld, n, n, update, Su);
  gemv_rows_kernel<<<rb, 256, 0, s>>>(S0, ld, n, n, update, S0u);
  solve_scalars_kernel<<<1, 1024, 0, s>>>(Su, S0u, F, update, n, meanE2, scalars);
  VMC_LAUNCH_CHECK("solve_tail");
  return 0;
}

// Solve S x = F by blocked Cholesky.  S (n x n, ld) is overwritten by its lower factor.  info (device int, must be
// zero on entry) receives 1 + index of the first non-positive pivot; the caller checks it after synchronising.
// scalars[2] (optional) = {residual ||S x - F||/||F||, tdvp_error} need S_copy / S0 / meanE2 like solve_tail.
extern "C" __attribute__((visibility("default"))) int vmcpde_chol_solve(double* S, int32_t n, int32_t ld, const double* F,
                                                                       double* x, int32_t* info, vmcpde_stream stream) {
  VMC_REQUIRE(S && F && x && info, "vmcpde_chol_solve: null pointer");
  VMC_REQUIRE(n >= 1 && ld >= n, "vmcpde_chol_solve: bad dimensions");
  cudaStream_t s = (cudaStream_t)stream;
  zero_diag_to_one_kernel<<<(n + 255) / 256, 256, 0, s>>>(S, ld, n);
  if (ld % 128 == 0 && n >= 256 && ld >= (n + 127) / 128 * 128 && !getenv("VMCPDE_CHOL_LEGACY")) {
    // tensor-core path (S is an ld x ld buffer, as every caller of this package allocates it; its lower triangle is scratch)
    const int np = (n + 127) / 128 * 128;
    const size_t psm = (size_t)kCb * (kCb + 1) * 8;
    VMC_CUDA_CHECK(cudaFuncSetAttribute(trsm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
    VMC_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(psm + kCb * 8)));
    VMC_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(psm + kCb * 8)));
    const bool timing = getenv("VMCPDE_CHOL_TIMING") != nullptr;   // debugging aid: synchronises
    float t_potrf = 0.f, t_trsm = 0.f, t_syrk = 0.f, t_sub = 0.f;
    cudaEvent_t ev[2];
    if (timing) { cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]); }
    auto tic = [&]() { if (timing) cudaEventRecord(ev[0], s); };
    auto toc = [&](float& acc) { if (timing) { cudaEventRecord(ev[1], s); cudaEventSynchronize(ev[1]); float ms; cudaEventElapsedTime(&ms, ev[0], ev[1]); acc += ms; } };
    for (int j0 = 0; j0 < np; j0 += kCb) {
      const int j1 = j0 + kCb, m = np - j1;
      tic(); potrf128_kernel<<<1, 256, 0, s>>>(S, ld, j0, n, info); toc(t_potrf);
      if (m > 0) {
        tic(); trsm128_kernel<<<(m + 127) / 128, 128, psm, s>>>(S, ld, j0, m); toc(t_trsm);
        double* A12 = S + (size_t)j0 * ld + j1;
        tic();
        if (int rc = vmcpde_syrk_tn(A12, ld, S + (size_t)j1 * ld + j1, ld, m, kCb, -1.0, 1.0, stream)) return rc;
        toc(t_syrk);
      }
    }
    tic();
    VMC_CUDA_CHECK(cudaMemcpyAsync(x, F, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    for (int j0 = 0; j0 < n; j0 += kCb) {
      chol_diag_forward_kernel<<<1, 128, psm + kCb * 8, s>>>(S, ld, n, j0, x);
      const int rest = n - j0 - kCb;
      if (rest > 0) chol_offblock_forward_kernel<<<(rest + 255) / 256, 256, 0, s>>>(S, ld, n, j0, x);
    }
    for (int j0 = ((n - 1) / kCb) * kCb; j0 >= 0; j0 -= kCb) {
      if (n - j0 - kCb > 0) chol_offblock_backward_kernel<<<kCb / 8, 256, 0, s>>>(S, ld, n, j0, x);
      chol_diag_backward_kernel<<<1, 128, psm + kCb * 8, s>>>(S, ld, n, j0, x);
    }
    toc(t_sub);
    if (timing) {
      fprintf(stderr, "[vmcpde_chol_solve n=%d] potrf %.2f ms, trsm %.2f ms, syrk %.2f ms, substitution %.2f ms\n", n, t_potrf, t_trsm, t_syrk, t_sub);
      cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    }
    VMC_LAUNCH_CHECK("chol_solve (tensor-core path)");
    return 0;
  }
  for (int j0 = 0; j0 < n; j0 += kNb) {
    const int nb = min(kNb, n - j0);
    potrf_diag_kernel<<<1, 256, 0, s>>>(S, ld, j0, nb, info);
    const int rem = n - j0 - nb;
    if (rem > 0) {
      trsm_panel_kernel<<<(rem + 63) / 64, 64, 0, s>>>(S, ld, n, j0, nb);
      const int tiles = (rem + 63) / 64;
      syrk_update_kernel<<<dim3(tiles, tiles), 128, 0, s>>>(S, ld, n, j0, nb);
    }
  }
  chol_substitute_kernel<<<1, 1024, 0, s>>>(S, ld, n, F, x);
  VMC_LAUNCH_CHECK("chol_solve");
  return 0;
}

// residual and TDVP error for a given update (used by the Cholesky path): scalars = {||S u - F||/||F||, tdvp_error}
extern "C" __attribute__((visibility("default"))) int vmcpde_solve_scalars(const double* S, const double* S0, int32_t n, int32_t ld,
                                                                          const double* F, const double* update, double meanE2,
                                                                          double* scalars, double* work2n, vmcpde_stream stream) {
  VMC_REQUIRE(S && S0 && F && update && scalars && work2n, "vmcpde_solve_scalars: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int rb = max(1, min(num_sms() * 4, (n + 7) / 8));
  gemv_rows_kernel<<<rb, 256, 0, s>>>(S, ld, n, n, update, work2n);
  gemv_rows_kernel<<<rb, 256, 0, s>>>(S0, ld, n, n, update, work2n + n);
  solve_scalars_kernel<<<1, 1024, 0, s>>>(work2n, work2n + n, F, update, n, meanE2, scalars);
  VMC_LAUNCH_CHECK("solve_scalars");
  return 0;
}
