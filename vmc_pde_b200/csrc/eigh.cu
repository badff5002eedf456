// Symmetric eigensolver on the device: S = V diag(ev) V^T, ev ascending.
// Replaces the host LAPACK round trip of the reference (np.array(S) -> np.linalg.eigh -> jnp.array,
// tdvp.py:57-64) with three device stages and no host synchronisation:
//   1. Householder tridiagonalisation (reflectors kept in the rows of the work matrix),
//   2. Cuppen divide & conquer on the tridiagonal (leaves of size 1; per level: deflation scan, Givens
//      deflation, secular roots, Loewner vector, eigenvector update as a DMMA GEMM); scalar core in
//      dc_core.cuh (validated on the host against LAPACK),
//   3. back-transformation: every eigenvector is a row that stays in shared memory while all
//      reflectors are applied to it (rows are independent, so one launch does the whole sweep).
// Eigenvectors are produced as rows (VT); the Python layer exposes V = VT^T as a view.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "dc_core.cuh"

namespace vmc {

// ================================================================================================
// Stage 0: scaling (the role of dlascl in LAPACK's dsyevd / dstedc): S is multiplied by the power of two that brings its
// largest entry into [0.5, 1) and the eigenvalues are scaled back at the end.  It keeps the squared norms of the
// Householder step in range and gives the divide & conquer deflation test a matrix of norm O(1).
// Everything stays on the device: sc[0] = max |S|, sc[1] = factor, sc[2] = 1 / factor.
// ================================================================================================
__global__ void __launch_bounds__(256) absmax_kernel(const double* __restrict__ S, int n, int ld, double* __restrict__ sc) {
  double m = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * ld; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    if (c < n) m = fmax(m, fabs(S[i]));   // NaN entries are ignored here and surface in the result
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0)
    atomicMax((unsigned long long*)sc, (unsigned long long)__double_as_longlong(m));  // monotone for non-negative doubles
}
__global__ void scale_factor_kernel(double* __restrict__ sc) {
  // max |S| * factor in [0.5, 1): exact (power of two), and the tridiagonal matrix enters divide & conquer with norm O(1)
  // like LAPACK's dstedc -- the deflation tolerance 8 eps max(|d|, |z|) compares eigenvalues with components of unit
  // vectors and over-deflates a matrix of small norm otherwise.
  const double anrm = sc[0];
  double f = 1.0;
  if (anrm > 0.0 && anrm < 1.7e308) {
    int e = 0;
    frexp(anrm, &e);
    e = max(-1000, min(1000, e));
    f = scalbn(1.0, -e);
  }
  sc[1] = f;
  sc[2] = 1.0 / f;
}
__global__ void __launch_bounds__(256) scale_by_kernel(double* __restrict__ x, size_t count, const double* __restrict__ factor) {
  const double f = *factor;
  if (f == 1.0) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) x[i] *= f;
}

// ================================================================================================
// Stage 1: tridiagonalisation (full symmetric storage, row-major, leading dimension ld)
// ================================================================================================
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result broadcast to all threads; `sh` holds >= 33 doubles
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (lane < nw) ? sh[lane] : 0.0;
  r = warp_sum(r);
  return r;
}

// row j: Householder vector of x = A[j][j+1..n); v (v[0] = 1) overwrites x; d[j], e[j], tau[j]
__global__ void __launch_bounds__(1024) tridiag_house_kernel(double* __restrict__ A, int ld, int n, int j,
                                                            double* __restrict__ d, double* __restrict__ e,
                                                            double* __restrict__ tau) {
  __shared__ double sh[33];
  double* x = A + (size_t)j * ld + j + 1;
  const int m = n - j - 1;
  double s = 0.0;
  for (int i = 1 + threadIdx.x; i < m; i += blockDim.x) s += x[i] * x[i];
  const double sigma = block_sum(s, sh);
  const double alpha = x[0];
  __syncthreads();
  if (sigma == 0.0) {
    if (threadIdx.x == 0) { d[j] = A[(size_t)j * ld + j]; e[j] = alpha; tau[j] = 0.0; x[0] = 1.0; }
    return;
  }
  const double beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
  const double scale = 1.0 / (alpha - beta);
  for (int i = 1 + threadIdx.x; i < m; i += blockDim.x) x[i] *= scale;
  if (threadIdx.x == 0) {
    d[j] = A[(size_t)j * ld + j]; e[j] = beta; tau[j] = (beta - alpha) / beta; x[0] = 1.0;
  }
}

// p[r] = sum_c A22[r][c] v[c], A22 = A[j+1.., j+1..], warp per row
__global__ void __launch_bounds__(256) tridiag_symv_kernel(const double* __restrict__ A, int ld, int n, int j,
                                                           double* __restrict__ p) {
  const int m = n - j - 1;
  const double* v = A + (size_t)j * ld + j + 1;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < m; r += gridDim.x * wpb) {
    const double* row = A + (size_t)(j + 1 + r) * ld + j + 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = lane;
    for (; c + 96 < m; c += 128) {
      const double a0 = row[c], a1 = row[c + 32], a2 = row[c + 64], a3 = row[c + 96];
      s0 = fma(a0, v[c], s0); s1 = fma(a1, v[c + 32], s1); s2 = fma(a2, v[c + 64], s2); s3 = fma(a3, v[c + 96], s3);
    }
    for (; c < m; c += 32) s0 = fma(row[c], v[c], s0);
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) p[r] = s;
  }
}

// w = tau*p - (tau^2/2)(p.v) v   (dsytd2: p' = tau A v; w = p' - (tau/2)(p'.v) v), stored over p
__global__ void __launch_bounds__(1024) tridiag_w_kernel(const double* __restrict__ A, int ld, int n, int j,
                                                         const double* __restrict__ tau, double* __restrict__ p) {
  __shared__ double sh[33];
  const int m = n - j - 1;
  const double* v = A + (size_t)j * ld + j + 1;
  const double t = tau[j];
  double s = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) s += p[i] * v[i];
  const double pv = block_sum(s, sh);
  const double alpha = -0.5 * t * t * pv;
  for (int i = threadIdx.x; i < m; i += blockDim.x) p[i] = t * p[i] + alpha * v[i];
}

// ---- fused step: the pending rank-2 update of step j is applied while the next Householder vector is formed
// and multiplied, so the trailing matrix is read and written once per column (16 B / element instead of 24).

// first row of A22(j) (global row j+1): apply the pending update, then form the next reflector from it
__global__ void __launch_bounds__(1024) tridiag_next_house_kernel(double* __restrict__ A, int ld, int n, int j,
                                                                 const double* __restrict__ w, double* __restrict__ d,
                                                                 double* __restrict__ e, double* __restrict__ tau) {
  __shared__ double sh[33];
  const int m = n - j - 1;                       // size of A22(j)
  const double* v = A + (size_t)j * ld + j + 1;  // v_j
  double* row = A + (size_t)(j + 1) * ld + j + 1;
  const double v0 = v[0], w0 = w[0];
  double s = 0.0;
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    const double a = row[c] - (v0 * w[c] + w0 * v[c]);
    row[c] = a;
    if (c >= 2) s += a * a;
  }
  const double sigma = block_sum(s, sh);
  __syncthreads();
  // x = row[1..m): alpha = row[1], rest from c = 2
  const double alpha = row[1];
  double* x = row + 1;
  const int mx = m - 1;
  if (sigma == 0.0) {
    if (threadIdx.x == 0) { d[j + 1] = row[0]; e[j + 1] = alpha; tau[j + 1] = 0.0; x[0] = 1.0; }
    return;
  }
  const double beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
  const double scale = 1.0 / (alpha - beta);
  for (int i = 1 + threadIdx.x; i < mx; i += blockDim.x) x[i] *= scale;
  if (threadIdx.x == 0) { d[j + 1] = row[0]; e[j + 1] = beta; tau[j + 1] = (beta - alpha) / beta; x[0] = 1.0; }
}

// rows r >= 1, cols c >= 1 of A22(j): a' = a - v_j[r] w_j[c] - w_j[r] v_j[c]; p'[r-1] = sum_c a' * v_{j+1}[c-1]
__global__ void __launch_bounds__(256) tridiag_update_symv_kernel(double* __restrict__ A, int ld, int n, int j,
                                                                 const double* __restrict__ w, double* __restrict__ pn) {
  const int m = n - j - 1;
  const double* v = A + (size_t)j * ld + j + 1;         // v_j, indices 0..m-1
  const double* vn = A + (size_t)(j + 1) * ld + j + 2;  // v_{j+1}, indices 0..m-2  (column c <-> vn[c-1])
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = 1 + blockIdx.x * wpb + (threadIdx.x >> 5); r < m; r += gridDim.x * wpb) {
    double* row = A + (size_t)(j + 1 + r) * ld + j + 1;
    const double vr = v[r], wr = w[r];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 1 + lane;
    for (; c + 96 < m; c += 128) {
      double a0 = row[c], a1 = row[c + 32], a2 = row[c + 64], a3 = row[c + 96];
      a0 -= vr * w[c] + wr * v[c];
      a1 -= vr * w[c + 32] + wr * v[c + 32];
      a2 -= vr * w[c + 64] + wr * v[c + 64];
      a3 -= vr * w[c + 96] + wr * v[c + 96];
      row[c] = a0; row[c + 32] = a1; row[c + 64] = a2; row[c + 96] = a3;
      s0 = fma(a0, vn[c - 1], s0); s1 = fma(a1, vn[c + 31], s1); s2 = fma(a2, vn[c + 63], s2); s3 = fma(a3, vn[c + 95], s3);
    }
    for (; c < m; c += 32) {
      const double a0 = row[c] - (vr * w[c] + wr * v[c]);
      row[c] = a0;
      s0 = fma(a0, vn[c - 1], s0);
    }
    const double sum = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) pn[r - 1] = sum;
  }
}

// w_{j+1} from p' (tridiag_w_kernel with the next reflector); separate name for clarity of the launch sequence
// last pending update on the final 2x2 block
__global__ void tridiag_last_update_kernel(double* __restrict__ A, int ld, int n, int j, const double* __restrict__ w) {
  const int m = n - j - 1;
  const double* v = A + (size_t)j * ld + j + 1;
  const int r = threadIdx.x / m, c = threadIdx.x % m;
  if (r < m) A[(size_t)(j + 1 + r) * ld + j + 1 + c] -= v[r] * w[c] + w[r] * v[c];
}

__global__ void tridiag_tail_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ d,
                                    double* __restrict__ e) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (n >= 2) { d[n - 2] = A[(size_t)(n - 2) * ld + n - 2]; e[n - 2] = A[(size_t)(n - 2) * ld + n - 1]; }
    d[n - 1] = A[(size_t)(n - 1) * ld + n - 1];
    e[n - 1] = 0.0;
  }
}

// ================================================================================================
// Stage 2: divide & conquer.  Per-node scratch lives in arrays of length n indexed by the node offset lo.
// ================================================================================================
struct DcBuf {
  int n, ld;
  const double* e;      // off-diagonals of the tridiagonal
  double *lam, *lam_new;
  double *QT, *QT_new, *U;
  double *z, *dl, *w, *w2, *dfv, *tau, *what, *vals, *norm;
  int *order, *nd, *dfi, *org, *pos;
  DcRot* rots;
  int *k, *nrot;        // per node (indexed by node index within the level)
  double* rho;          // per node
};

__global__ void dc_init_kernel(DcBuf b, const double* __restrict__ d) {
  const int n = b.n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * b.ld; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / b.ld), c = (int)(i % b.ld);
    b.QT[i] = (r == c && c < n) ? 1.0 : 0.0;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    b.lam[i] = d[i] - (i > 0 ? fabs(b.e[i - 1]) : 0.0) - (i + 1 < n ? fabs(b.e[i]) : 0.0);
}

constexpr int kDeflateSmemMax = 11000;  // node sizes up to this stage (dmod, z, order) in shared memory

// one CTA per node: z gather, deflation scan (thread 0), w^2
__global__ void __launch_bounds__(256) dc_deflate_kernel(DcBuf b, int depth) {
  extern __shared__ double sm[];
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  const int nm = hi - lo, mid = lo + nm / 2, n1 = mid - lo;
  if (nm <= 1 || n1 == 0) {  // nothing to merge: carry the eigenpair over
    if (threadIdx.x == 0) {
      b.k[blockIdx.x] = -1;
      for (int r = lo; r < hi; ++r) {
        b.lam_new[r] = b.lam[r];
        for (int c = lo; c < hi; ++c) b.QT_new[(size_t)r * b.ld + c] = b.QT[(size_t)r * b.ld + c];
      }
    }
    return;
  }
  const double rs = b.e[mid - 1];
  const bool staged = nm <= kDeflateSmemMax;
  double* dmod = staged ? sm : b.vals + lo;            // vals is free at this point of the level
  double* z = staged ? sm + nm : b.z + lo;
  int* order = staged ? (int*)(sm + 2 * nm) : b.order + lo;
  for (int c = threadIdx.x; c < nm; c += blockDim.x) {
    z[c] = c < n1 ? b.QT[(size_t)(lo + c) * b.ld + mid - 1] : (rs < 0.0 ? -1.0 : 1.0) * b.QT[(size_t)(lo + c) * b.ld + mid];
    dmod[c] = b.lam[lo + c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int k, nrot;
    double rho;
    dc_deflate(nm, n1, rs, dmod, z, dmod, order, b.dl + lo, b.w + lo, b.nd + lo, b.dfv + lo, b.dfi + lo, b.rots + lo,
               &k, &nrot, &rho);
    b.k[blockIdx.x] = k; b.nrot[blockIdx.x] = nrot; b.rho[blockIdx.x] = rho;
  }
  __syncthreads();
  const int k = b.k[blockIdx.x];
  for (int i = threadIdx.x; i < k; i += blockDim.x) { const double wv = b.w[lo + i]; b.w2[lo + i] = wv * wv; }
}

// Givens deflation applied to the eigenvector rows of the node (columns lo..hi)
__global__ void __launch_bounds__(256) dc_rotate_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  if (b.k[blockIdx.x] < 0) return;
  const int nrot = b.nrot[blockIdx.x];
  for (int r = 0; r < nrot; ++r) {
    const DcRot q = b.rots[lo + r];
    double* x = b.QT + (size_t)(lo + q.a) * b.ld;
    double* y = b.QT + (size_t)(lo + q.b) * b.ld;
    for (int c = lo + blockIdx.y * blockDim.x + threadIdx.x; c < hi; c += gridDim.y * blockDim.x) {
      const double xv = x[c], yv = y[c];
      x[c] = q.c * xv + q.s * yv;
      y[c] = q.c * yv - q.s * xv;
    }
  }
}

struct WarpSums {  // warp-cooperative secular sums; every lane receives the totals
  const double* dl;
  const double* w2;
  int k;
  __host__ __device__ SecularSums operator()(int o, double tau, int jl) const {
    SecularSums r{0, 0, 0, 0, 0};
#ifdef __CUDA_ARCH__
    const double dlo = dl[o];
    for (int i = threadIdx.x & 31; i < k; i += 32) {
      const double del = (dl[i] - dlo) - tau;
      const double t = w2[i] / del;
      if (i <= jl) { r.psi += t; r.dpsi += t / del; } else { r.phi += t; r.dphi += t / del; }
      r.sabs += fabs(t);
    }
    r.psi = warp_sum(r.psi); r.dpsi = warp_sum(r.dpsi); r.phi = warp_sum(r.phi); r.dphi = warp_sum(r.dphi);
    r.sabs = warp_sum(r.sabs);
#endif
    return r;
  }
};

// warp per root
__global__ void __launch_bounds__(256) dc_secular_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  const int k = b.k[blockIdx.x];
  if (k <= 0) return;
  const int nm = hi - lo;
  const double rho = b.rho[blockIdx.x];
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  WarpSums ws{b.dl + lo, b.w2 + lo, k};
  for (int j = blockIdx.y * wpb + (threadIdx.x >> 5); j < nm; j += gridDim.y * wpb) {
    if (j < k) {
      int o;
      double t;
      secular_root(k, j, b.dl + lo, b.w2 + lo, rho, ws, &o, &t);
      if (lane == 0) { b.org[lo + j] = o; b.tau[lo + j] = t; b.vals[lo + j] = b.dl[lo + o] + t; }
    } else if (lane == 0) {
      b.vals[lo + j] = b.dfv[lo + j - k];
    }
  }
}
__global__ void __launch_bounds__(256) dc_alldeflated_vals_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  if (b.k[blockIdx.x] != 0) return;
  for (int j = threadIdx.x; j < hi - lo; j += blockDim.x) b.vals[lo + j] = b.dfv[lo + j];
}

// rank of every new eigenvalue inside the node -> pos ; lam_new
__global__ void __launch_bounds__(256) dc_rank_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  if (b.k[blockIdx.x] < 0) return;
  const int nm = hi - lo;
  const double* v = b.vals + lo;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < nm; i += gridDim.y * blockDim.x) {
    const double vi = v[i];
    int r = 0;
    for (int q = 0; q < nm; ++q) { const double vq = v[q]; r += (vq < vi) || (vq == vi && q < i); }
    b.pos[lo + i] = r;
    b.lam_new[lo + r] = vi;
  }
}

__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Loewner vector, warp per component i
__global__ void __launch_bounds__(256) dc_lowner_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  const int k = b.k[blockIdx.x];
  if (k <= 0) return;
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  const double* dl = b.dl + lo;
  const int* org = b.org + lo;
  const double* tau = b.tau + lo;
  for (int i = blockIdx.y * wpb + (threadIdx.x >> 5); i < k; i += gridDim.y * wpb) {
    double p = 1.0;
    const double dli = dl[i];
    for (int j = lane; j < k; j += 32) {
      const double del = (dli - dl[org[j]]) - tau[j];
      p *= (j == i) ? del : del / (dli - dl[j]);
    }
    p = warp_prod(p);
    if (lane == 0) { const double v = sqrt(fabs(p)); b.what[lo + i] = b.w[lo + i] < 0.0 ? -v : v; }
  }
}

// column norms of what_i / delta(i,j), warp per column j
__global__ void __launch_bounds__(256) dc_colnorm_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  const int k = b.k[blockIdx.x];
  if (k <= 0) return;
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  const double* dl = b.dl + lo;
  for (int j = blockIdx.y * wpb + (threadIdx.x >> 5); j < k; j += gridDim.y * wpb) {
    const double dlo = dl[b.org[lo + j]], t = b.tau[lo + j];
    double s = 0.0;
    for (int i = lane; i < k; i += 32) { const double u = b.what[lo + i] / ((dl[i] - dlo) - t); s += u * u; }
    s = warp_sum(s);
    if (lane == 0) b.norm[lo + j] = sqrt(s);
  }
}

// U[i][j] = what_i / (delta(i,j) * norm_j), stored at U[(lo+i)*ld + lo + j]; rows/cols padded with zeros up to kp
__global__ void __launch_bounds__(256) dc_umat_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  const int k = b.k[blockIdx.x];
  if (k <= 0) return;
  const double* dl = b.dl + lo;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = blockIdx.y * 8 + ty; i < k; i += gridDim.y * 8) {
    const double wi = b.what[lo + i], dli = dl[i];
    double* urow = b.U + (size_t)(lo + i) * b.ld + lo;
    for (int j = tx; j < k; j += 32) urow[j] = wi / (((dli - dl[b.org[lo + j]]) - b.tau[lo + j]) * b.norm[lo + j]);
  }
}

// QT_new[lo + pos[j]][lo + c] = sum_i U[lo+i][lo+j] * QT[lo + nd[i]][lo + c]   (j < k, c < nm, i < k)
// DMMA tiles 64 (j) x 64 (c), K chunks of 16, operands staged through padded shared memory.
constexpr int kGT = 64, kGK = 16, kGLd = 68;
__device__ __forceinline__ void dmma_e(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) dc_gemm_kernel(DcBuf b, int depth) {
  __shared__ double Xs[kGK][kGLd], Ys[kGK][kGLd];
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.z, lo, hi);
  const int k = b.k[blockIdx.z];
  if (k <= 0) return;
  const int nm = hi - lo;
  const int j0 = blockIdx.y * kGT, c0 = blockIdx.x * kGT;
  if (j0 >= k || c0 >= nm) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;  // 2x2 warps, each 32 x 32
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
  const int* nd = b.nd + lo;
  for (int i0 = 0; i0 < k; i0 += kGK) {
    // stage X = U rows i0.., cols j0..j0+64 ; Y = QT rows nd[i0..], cols c0..c0+64
    for (int idx = threadIdx.x; idx < kGK * kGT; idx += 128) {
      const int r = idx >> 6, c = idx & 63;
      const int i = i0 + r;
      double xv = 0.0, yv = 0.0;
      if (i < k) {
        if (j0 + c < k) xv = b.U[(size_t)(lo + i) * b.ld + lo + j0 + c];
        if (c0 + c < nm) yv = b.QT[(size_t)(lo + nd[i]) * b.ld + lo + c0 + c];
      }
      Xs[r][c] = xv; Ys[r][c] = yv;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kGK / 4; ++s) {
      double a[4], bb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = Xs[s * 4 + t][wm * 32 + q * 8 + g]; bb[q] = Ys[s * 4 + t][wn * 32 + q * 8 + g]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p) dmma_e(acc[q][p][0], acc[q][p][1], a[q], bb[p]);
    }
    __syncthreads();
  }
  const int* pos = b.pos + lo;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = j0 + wm * 32 + q * 8 + g;
    if (j >= k) continue;
    double* out = b.QT_new + (size_t)(lo + pos[j]) * b.ld + lo;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int c = c0 + wn * 32 + p * 8 + 2 * t;
      if (c < nm) out[c] = acc[q][p][0];
      if (c + 1 < nm) out[c + 1] = acc[q][p][1];
    }
  }
}

// deflated eigenvectors are copied to their sorted slot
__global__ void __launch_bounds__(256) dc_copy_deflated_kernel(DcBuf b, int depth) {
  int lo, hi;
  dc_node_range(b.n, depth, blockIdx.x, lo, hi);
  const int k = b.k[blockIdx.x];
  if (k < 0) return;
  const int nm = hi - lo;
  for (int m = k + blockIdx.y; m < nm; m += gridDim.y) {
    const double* src = b.QT + (size_t)(lo + b.dfi[lo + m - k]) * b.ld + lo;
    double* dst = b.QT_new + (size_t)(lo + b.pos[lo + m]) * b.ld + lo;
    for (int c = threadIdx.x; c < nm; c += blockDim.x) dst[c] = src[c];
  }
}

// ================================================================================================
// Stage 3: back-transformation  VT[k][:] = ZT[k][:] * H_{n-3} ... H_0   (rows independent)
// ================================================================================================
// Generic version (any n <= 25600): rows live in shared memory.
__global__ void __launch_bounds__(256) backtransform_smem_kernel(const double* __restrict__ ZT, double* __restrict__ VT,
                                                            const double* __restrict__ A, const double* __restrict__ tau,
                                                            int n, int ld, int rows_per_cta) {
  extern __shared__ double rows[];  // rows_per_cta * n, then 2 * 8 * rows_per_cta partials
  double* part = rows + (size_t)rows_per_cta * n;
  const int r0 = blockIdx.x * rows_per_cta;
  const int nr = min(rows_per_cta, n - r0);
  if (nr <= 0) return;
  for (int r = 0; r < nr; ++r)
    for (int c = threadIdx.x; c < n; c += blockDim.x) rows[(size_t)r * n + c] = ZT[(size_t)(r0 + r) * ld + c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int par = 0;
  for (int j = n - 3; j >= 0; --j) {
    const double t = tau[j];
    if (t == 0.0) continue;
    const double* v = A + (size_t)j * ld;  // v[c] valid for c in [j+1, n)
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    // each thread always owns columns c = tid (mod 256): no cross-thread hazards on the rows
    const int cstart = (int)threadIdx.x > j ? (int)threadIdx.x
                                            : (int)threadIdx.x + ((j + 1 - (int)threadIdx.x + 255) / 256) * 256;
    for (int c = cstart; c < n; c += 256) {
      const double vc = v[c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < nr) s[r] = fma(vc, rows[(size_t)r * n + c], s[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < nr) {
        const double ws = warp_sum(s[r]);
        if (lane == 0) part[(par * 8 + warp) * 4 + r] = ws;
      }
    }
    __syncthreads();
    double tot[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      tot[r] = 0.0;
      if (r < nr) {
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tot[r] += part[(par * 8 + w8) * 4 + r];
        tot[r] *= t;
      }
    }
    for (int c = cstart; c < n; c += 256) {
      const double vc = v[c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < nr) rows[(size_t)r * n + c] = fma(-tot[r], vc, rows[(size_t)r * n + c]);
    }
    par ^= 1;
    // the next iteration's partials go to the other buffer; its __syncthreads orders the row updates
  }
  __syncthreads();
  for (int r = 0; r < nr; ++r)
    for (int c = threadIdx.x; c < n; c += blockDim.x) VT[(size_t)(r0 + r) * ld + c] = rows[(size_t)r * n + c];
}


// Fast version for n <= 8192: each thread owns columns c = tid + k*1024 (k < CPT) of ROWS eigenvector rows and keeps
// them in registers for the whole sweep; reflectors stream through a double-buffered shared-memory stage filled with
// cp.async by the threads that will consume them (no cross-thread hand-off).
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <int CPT, int ROWS>
__global__ void __launch_bounds__(1024, 1) backtransform_reg_kernel(const double* __restrict__ ZT, double* __restrict__ VT,
                                                                    const double* __restrict__ A, const double* __restrict__ tau,
                                                                    int n, int ld) {
  extern __shared__ double vstage[];  // [2][CPT * 1024]
  __shared__ double part[2][32][ROWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * ROWS;
  double x[ROWS][CPT];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * 1024;
      x[r][k] = (r0 + r < n && c < n) ? ZT[(size_t)(r0 + r) * ld + c] : 0.0;
    }
  auto prefetch = [&](int j, int buf) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * 1024;
      if (c > j && c < n) cp_async8(&vstage[buf * CPT * 1024 + c], A + (size_t)j * ld + c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n >= 3) prefetch(n - 3, 0);
  int par = 0, buf = 0;
  for (int j = n - 3; j >= 0; --j) {
    if (j > 0) prefetch(j - 1, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const double t = tau[j];
    const double* vs = vstage + buf * CPT * 1024;
    buf ^= 1;
    if (t == 0.0) continue;
    double v[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) { const int c = tid + k * 1024; v[k] = (c > j && c < n) ? vs[c] : 0.0; }
    double sdot[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < CPT; ++k) acc = fma(v[k], x[r][k], acc);
      sdot[r] = warp_sum(acc);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) part[par][warp][r] = sdot[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const double tot = t * warp_sum(part[par][lane][r]);
#pragma unroll
      for (int k = 0; k < CPT; ++k) x[r][k] = fma(-tot, v[k], x[r][k]);
    }
    par ^= 1;
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * 1024;
      if (r0 + r < n && c < n) VT[(size_t)(r0 + r) * ld + c] = x[r][k];
    }
}

template <int CPT, int ROWS>
static int launch_backtransform(const double* ZT, double* VT, const double* A, const double* tau, int n, int ld, cudaStream_t s) {
  const size_t smem = (size_t)2 * CPT * 1024 * 8;
  VMC_CUDA_CHECK(cudaFuncSetAttribute(backtransform_reg_kernel<CPT, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  backtransform_reg_kernel<CPT, ROWS><<<(n + ROWS - 1) / ROWS, 1024, smem, s>>>(ZT, VT, A, tau, n, ld);
  return 0;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// eigh_blocked.cu
bool blocked_eigh_supported(int n, int ld);
size_t blocked_tridiag_scratch_bytes(int n, int ld);
size_t blocked_backtransform_scratch_bytes(int n, int ld);
int blocked_tridiag_launches(int n);
int blocked_backtransform_launches(int n);
int tridiag_blocked(double* A, int n, int ld, double* d, double* e, double* tau, void* scratch, size_t scratch_bytes,
                    cudaStream_t s);
int backtransform_blocked(double* A, const double* tau, int n, int ld, const double* ZT, double* Zn, double* AT,
                          void* scratch, size_t scratch_bytes, double* VT, int col0, int ncols, cudaStream_t s);

}  // namespace vmc

using namespace vmc;

extern "C" __attribute__((visibility("default"))) int vmcpde_eigh_workspace_bytes(int32_t n, int32_t ld, size_t* bytes) {
  VMC_REQUIRE(bytes && n >= 1 && ld >= n, "vmcpde_eigh_workspace_bytes: bad arguments");
  size_t b = 0;
  const bool blocked = blocked_eigh_supported(n, ld);
  const size_t rows = blocked ? (size_t)(n + 127) / 128 * 128 : (size_t)n;
  b += 2 * align_up(rows * ld * 8, 256);                // QT ping buffer, U
  if (blocked) b += align_up(blocked_tridiag_scratch_bytes(n, ld), 256) + align_up(blocked_backtransform_scratch_bytes(n, ld), 256);
  b += 18 * align_up((size_t)(n + 8) * 8, 256);          // double vectors
  b += 8 * align_up((size_t)(n + 8) * 4, 256);           // int vectors
  b += align_up((size_t)(n + 8) * sizeof(DcRot), 256);  // rotations
  *bytes = b;
  return 0;
}

// Kernel launches issued by one vmcpde_eigh call (the bench reports it in gpu_launches).
extern "C" __attribute__((visibility("default"))) int vmcpde_eigh_launch_count(int32_t n, int32_t ld, int32_t* count) {
  VMC_REQUIRE(count && n >= 1 && ld >= n, "vmcpde_eigh_launch_count: bad arguments");
  const int depth = dc_tree_depth(n);
  int c = 4 + 1 + 10 * depth;  // scaling (3 + 1), dc_init, per-level kernels
  if (blocked_eigh_supported(n, ld)) c += blocked_tridiag_launches(n) + blocked_backtransform_launches(n);
  else c += (n >= 3 ? 3 + 3 * (n - 3) + 1 : 0) + 1 + 1;
  *count = c;
  return 0;
}

// Workspace layout shared by the stages of the eigensolver (same carving order for every entry point, so a workspace
// of vmcpde_eigh_workspace_bytes serves any of them).
struct EighWs {
  bool blocked;
  size_t rows, tri_bytes, bt_bytes;
  double *QTb, *U, *d, *e, *tau, *p, *p2, *sc;
  void *tri_scratch, *bt_scratch;
  DcBuf b;
};

static int eigh_layout(EighWs& w, int32_t n, int32_t ld, void* workspace, size_t workspace_bytes) {
  VMC_REQUIRE(workspace, "vmcpde_eigh: null pointer");
  VMC_REQUIRE(n >= 1 && ld >= n, "vmcpde_eigh: bad dimensions");
  size_t need = 0;
  vmcpde_eigh_workspace_bytes(n, ld, &need);
  VMC_REQUIRE(workspace_bytes >= need, "vmcpde_eigh: workspace too small");
  VMC_REQUIRE(n <= 25 * 1024, "vmcpde_eigh: n > 25600 not supported in this release");
  uint8_t* wp = (uint8_t*)workspace;
  auto take = [&](size_t bytes) { void* q = wp; wp += align_up(bytes, 256); return q; };
  w.blocked = blocked_eigh_supported(n, ld);
  w.rows = w.blocked ? (size_t)(n + 127) / 128 * 128 : (size_t)n;
  w.QTb = (double*)take(w.rows * ld * 8);
  w.U = (double*)take(w.rows * ld * 8);
  w.tri_bytes = w.blocked ? blocked_tridiag_scratch_bytes(n, ld) : 0;
  w.bt_bytes = w.blocked ? blocked_backtransform_scratch_bytes(n, ld) : 0;
  w.tri_scratch = w.blocked ? take(w.tri_bytes) : nullptr;
  w.bt_scratch = w.blocked ? take(w.bt_bytes) : nullptr;
  auto dvec = [&]() { return (double*)take((size_t)(n + 8) * 8); };
  auto ivec = [&]() { return (int*)take((size_t)(n + 8) * 4); };
  w.d = dvec(); w.e = dvec(); w.tau = dvec(); w.p = dvec(); w.p2 = dvec();
  DcBuf& b = w.b;
  b = DcBuf{};
  b.n = n; b.ld = ld; b.e = w.e;
  b.lam = dvec(); b.lam_new = dvec();
  b.z = dvec(); b.dl = dvec(); b.w = dvec(); b.w2 = dvec(); b.dfv = dvec(); b.tau = dvec(); b.what = dvec();
  b.vals = dvec(); b.norm = dvec(); b.rho = dvec();
  b.order = ivec(); b.nd = ivec(); b.dfi = ivec(); b.org = ivec(); b.pos = ivec(); b.k = ivec(); b.nrot = ivec();
  b.rots = (DcRot*)take((size_t)(n + 8) * sizeof(DcRot));
  b.QT_new = w.QTb; b.U = w.U;
  w.sc = dvec();
  return 0;
}

// Stages 0-2: scaling, tridiagonalisation (S is overwritten by the reflectors, w.tau), divide & conquer.  `QT0` is the first
// of the two ping-pong buffers of the divide & conquer; on return w.b.QT holds Z^T (eigenvectors of the tridiagonal matrix as
// rows) and is either QT0 or w.QTb.
static int eigh_factor_impl(EighWs& w, double* S, int32_t n, int32_t ld, double* ev, double* QT0, cudaStream_t s,
                            cudaEvent_t* evt) {
  DcBuf& b = w.b;
  b.QT = QT0;
  const bool blocked = w.blocked;
  double *d = w.d, *e = w.e, *tau = w.tau, *p = w.p, *p2 = w.p2, *sc = w.sc;
  // ---- stage 0
  const int sms = num_sms();
  VMC_CUDA_CHECK(cudaMemsetAsync(sc, 0, 3 * sizeof(double), s));
  absmax_kernel<<<sms * 8, 256, 0, s>>>(S, n, ld, sc);
  scale_factor_kernel<<<1, 1, 0, s>>>(sc);
  scale_by_kernel<<<sms * 8, 256, 0, s>>>(S, (size_t)n * ld, sc + 1);
  // ---- stage 1
  if (blocked) {
    if (int rc = tridiag_blocked(S, n, ld, d, e, tau, w.tri_scratch, w.tri_bytes, s)) return rc;
  } else if (n >= 3) {
    double* wbuf[2] = {p, p2};
    {  // prologue: reflector 0 and its w
      const int m = n - 1;
      const int blocks = max(1, min(sms * 8, (m + 7) / 8));
      tridiag_house_kernel<<<1, 1024, 0, s>>>(S, ld, n, 0, d, e, tau);
      tridiag_symv_kernel<<<blocks, 256, 0, s>>>(S, ld, n, 0, wbuf[0]);
      tridiag_w_kernel<<<1, 1024, 0, s>>>(S, ld, n, 0, tau, wbuf[0]);
    }
    for (int j = 0; j + 2 < n; ++j) {
      const int m = n - j - 1;
      double* wj = wbuf[j & 1];
      double* wn = wbuf[(j + 1) & 1];
      if (j + 3 < n) {
        const int blocks = max(1, min(sms * 8, (m - 1 + 7) / 8));
        tridiag_next_house_kernel<<<1, 1024, 0, s>>>(S, ld, n, j, wj, d, e, tau);
        tridiag_update_symv_kernel<<<blocks, 256, 0, s>>>(S, ld, n, j, wj, wn);
        tridiag_w_kernel<<<1, 1024, 0, s>>>(S, ld, n, j + 1, tau, wn);
      } else {
        tridiag_last_update_kernel<<<1, 32, 0, s>>>(S, ld, n, j, wj);  // m == 2
      }
    }
  }
  if (!blocked) tridiag_tail_kernel<<<1, 32, 0, s>>>(S, ld, n, d, e);
  VMC_LAUNCH_CHECK("tridiagonalisation");

  if (evt) cudaEventRecord(evt[1], s);
  // ---- stage 2 (the ping buffer must be zero outside the diagonal blocks written level by level)
  VMC_CUDA_CHECK(cudaMemsetAsync(w.QTb, 0, w.rows * ld * 8, s));
  if (w.rows > (size_t)n) VMC_CUDA_CHECK(cudaMemsetAsync(QT0 + (size_t)n * ld, 0, (w.rows - n) * ld * 8, s));
  dc_init_kernel<<<sms * 4, 256, 0, s>>>(b, d);
  const int D = dc_tree_depth(n);
  static bool attr = false;
  if (!attr) {
    VMC_CUDA_CHECK(cudaFuncSetAttribute(dc_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDeflateSmemMax * 20 + 64));
    attr = true;
  }
  for (int depth = D - 1; depth >= 0; --depth) {
    const int nodes = 1 << depth;
    const int max_nm = (n + nodes - 1) / nodes;  // node sizes at one depth differ by at most one
    const size_t dsm = max_nm <= kDeflateSmemMax ? (size_t)max_nm * 20 + 64 : 0;
    const int ychunks = max(1, min(sms * 2 / nodes + 1, (max_nm + 7) / 8));
    dc_deflate_kernel<<<nodes, 256, dsm, s>>>(b, depth);
    dc_rotate_kernel<<<dim3(nodes, max(1, min(8, (max_nm + 255) / 256))), 256, 0, s>>>(b, depth);
    dc_alldeflated_vals_kernel<<<nodes, 256, 0, s>>>(b, depth);
    dc_secular_kernel<<<dim3(nodes, ychunks), 256, 0, s>>>(b, depth);
    dc_rank_kernel<<<dim3(nodes, max(1, min(sms * 2 / nodes + 1, (max_nm + 255) / 256))), 256, 0, s>>>(b, depth);
    dc_lowner_kernel<<<dim3(nodes, ychunks), 256, 0, s>>>(b, depth);
    dc_colnorm_kernel<<<dim3(nodes, ychunks), 256, 0, s>>>(b, depth);
    dc_umat_kernel<<<dim3(nodes, ychunks), 256, 0, s>>>(b, depth);
    const int tiles = (max_nm + kGT - 1) / kGT;
    dc_gemm_kernel<<<dim3(tiles, tiles, nodes), 128, 0, s>>>(b, depth);
    dc_copy_deflated_kernel<<<dim3(nodes, max(1, min(max_nm, sms * 4 / nodes + 1))), 256, 0, s>>>(b, depth);
    VMC_LAUNCH_CHECK("divide and conquer level");
    double* tq = b.QT; b.QT = b.QT_new; b.QT_new = tq;
    double* tl = b.lam; b.lam = b.lam_new; b.lam_new = tl;
  }
  VMC_CUDA_CHECK(cudaMemcpyAsync(ev, b.lam, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
  scale_by_kernel<<<max(1, min(sms, (n + 255) / 256)), 256, 0, s>>>(ev, (size_t)n, sc + 2);
  return 0;
}

// Stage 3: V = Q Z for the eigenvectors [col0, col0 + ncols).  A holds the reflectors (read only), ZT the rows of Z^T,
// AT is a free rows x ld buffer for the masked reflector transpose; VT may alias ZT or AT on the blocked path.
static int eigh_back_impl(EighWs& w, double* A, const double* tau, int32_t n, int32_t ld, const double* ZT, double* AT,
                          double* VT, int32_t col0, int32_t ncols, cudaStream_t s) {
  if (w.blocked) {
    const int np = (n + 127) / 128 * 128;
    if (ncols <= 0) { col0 = 0; ncols = np; }
    if (int rc = backtransform_blocked(A, tau, n, ld, ZT, w.U, AT, w.bt_scratch, w.bt_bytes, VT, col0, ncols, s)) return rc;
  } else {
    if (ZT == VT) {
      VMC_CUDA_CHECK(cudaMemcpyAsync(AT, VT, (size_t)n * ld * 8, cudaMemcpyDeviceToDevice, s));
      ZT = AT;
    }
    const int cpt = (n + 1023) / 1024;
    int rc = 0;
    if (cpt <= 1) rc = launch_backtransform<1, 8>(ZT, VT, A, tau, n, ld, s);
    else if (cpt <= 2) rc = launch_backtransform<2, 8>(ZT, VT, A, tau, n, ld, s);
    else if (cpt <= 4) rc = launch_backtransform<4, 4>(ZT, VT, A, tau, n, ld, s);
    else if (cpt <= 8) rc = launch_backtransform<8, 2>(ZT, VT, A, tau, n, ld, s);
    else {
      int rpc = (int)((200 * 1024 - 1024) / ((size_t)n * 8));
      if (rpc > 4) rpc = 4;
      if (rpc < 1) rpc = 1;
      const size_t bsm = (size_t)rpc * n * 8 + 2 * 8 * 4 * 8;
      VMC_CUDA_CHECK(cudaFuncSetAttribute(backtransform_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      backtransform_smem_kernel<<<(n + rpc - 1) / rpc, 256, bsm, s>>>(ZT, VT, A, tau, n, ld, rpc);
    }
    if (rc) return rc;
  }
  VMC_LAUNCH_CHECK("backtransform_kernel");
  return 0;
}

// S (n x n, leading dimension ld, full symmetric) is destroyed.  ev[n] ascending; VT row k = eigenvector k.
// Replaces np.linalg.eigh at tdvp.py:61-64.
static int eigh_impl(double* S, int32_t n, int32_t ld, double* ev, double* VT, int32_t col0, int32_t ncols, void* workspace,
                     size_t workspace_bytes, vmcpde_stream stream) {
  VMC_REQUIRE(S && ev && VT && workspace, "vmcpde_eigh: null pointer");
  EighWs w;
  if (int rc = eigh_layout(w, n, ld, workspace, workspace_bytes)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool timing = getenv("VMCPDE_EIGH_TIMING") != nullptr;
  cudaEvent_t evt[4];
  if (timing) { for (auto& e_ : evt) cudaEventCreate(&e_); cudaEventRecord(evt[0], s); }
  if (int rc = eigh_factor_impl(w, S, n, ld, ev, VT, s, timing ? evt : nullptr)) return rc;
  if (timing) cudaEventRecord(evt[2], s);
  // out of place: Z^T = b.QT -> VT; the other ping-pong buffer is free for the reflector transpose
  if (int rc = eigh_back_impl(w, S, w.tau, n, ld, w.b.QT, w.b.QT_new, VT, col0, ncols, s)) return rc;
  if (timing) {
    cudaEventRecord(evt[3], s);
    cudaEventSynchronize(evt[3]);
    float t1, t2, t3;
    cudaEventElapsedTime(&t1, evt[0], evt[1]); cudaEventElapsedTime(&t2, evt[1], evt[2]); cudaEventElapsedTime(&t3, evt[2], evt[3]);
    fprintf(stderr, "[vmcpde_eigh n=%d] tridiag %.2f ms, divide&conquer %.2f ms, backtransform %.2f ms\n", n, t1, t2, t3);
    for (auto& e_ : evt) cudaEventDestroy(e_);
  }
  return 0;
}

// The eigensolver in two calls, for a multi-GPU solve in which ONE rank factorises while the others still build Gram
// matrices: vmcpde_eigh_factor runs scaling, tridiagonalisation and divide & conquer (the serial part) and leaves the three
// things the back-transformation needs -- the reflectors (in S), tau[n] and Z^T (rows = eigenvectors of the tridiagonal
// matrix, rows x ld with rows = n rounded up to 128) -- in caller-owned buffers that can be broadcast;
// vmcpde_eigh_backtransform applies the reflectors to a 128-aligned slice of eigenvectors on any rank.
// factor + backtransform of all columns == vmcpde_eigh, bit for bit.  Blocked path only (n >= 384, padded ld).
extern "C" __attribute__((visibility("default"))) int vmcpde_eigh_factor(double* S, int32_t n, int32_t ld, double* ev, double* ZT,
                                                                        double* tau, void* workspace, size_t workspace_bytes,
                                                                        vmcpde_stream stream) {
  VMC_REQUIRE(S && ev && ZT && tau && workspace, "vmcpde_eigh_factor: null pointer");
  EighWs w;
  if (int rc = eigh_layout(w, n, ld, workspace, workspace_bytes)) return rc;
  if (!w.blocked) return set_error(VMCPDE_EUNSUPPORTED, "vmcpde_eigh_factor: needs the blocked path (n >= 384, ld a multiple of 128)");
  cudaStream_t s = (cudaStream_t)stream;
  // the divide & conquer only writes the n x n corner: the padding of Z^T must be zero for the back-transformation
  VMC_CUDA_CHECK(cudaMemsetAsync(ZT, 0, w.rows * ld * 8, s));
  if (int rc = eigh_factor_impl(w, S, n, ld, ev, ZT, s, nullptr)) return rc;
  if (w.b.QT != ZT) VMC_CUDA_CHECK(cudaMemcpyAsync(ZT, w.b.QT, w.rows * ld * 8, cudaMemcpyDeviceToDevice, s));
  VMC_CUDA_CHECK(cudaMemcpyAsync(tau, w.tau, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" __attribute__((visibility("default"))) int vmcpde_eigh_backtransform(const double* reflectors, const double* tau,
                                                                               const double* ZT, int32_t n, int32_t ld, double* VT,
                                                                               int32_t col0, int32_t ncols, void* workspace,
                                                                               size_t workspace_bytes, vmcpde_stream stream) {
  VMC_REQUIRE(reflectors && tau && ZT && VT && workspace, "vmcpde_eigh_backtransform: null pointer");
  VMC_REQUIRE(VT != ZT && VT != reflectors, "vmcpde_eigh_backtransform: VT must not alias the inputs");
  EighWs w;
  if (int rc = eigh_layout(w, n, ld, workspace, workspace_bytes)) return rc;
  if (!w.blocked) return set_error(VMCPDE_EUNSUPPORTED, "vmcpde_eigh_backtransform: needs the blocked path (n >= 384, ld a multiple of 128)");
  return eigh_back_impl(w, const_cast<double*>(reflectors), tau, n, ld, ZT, w.QTb, VT, col0, ncols, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int vmcpde_eigh(double* S, int32_t n, int32_t ld, double* ev, double* VT,
                                                                 void* workspace, size_t workspace_bytes,
                                                                 vmcpde_stream stream) {
  return eigh_impl(S, n, ld, ev, VT, 0, 0, workspace, workspace_bytes, stream);
}

// Same decomposition, but only the eigenvectors [col0, col0 + ncols) are back-transformed and written (rows col0.. of VT;
// the other rows are left untouched): the slice one rank of a multi-GPU solve needs.  col0 and ncols are multiples of 128
// inside the padded size when the blocked path runs (n >= 384, padded ld); the unblocked path writes every row.
extern "C" __attribute__((visibility("default"))) int vmcpde_eigh_cols(double* S, int32_t n, int32_t ld, double* ev, double* VT,
                                                                      int32_t col0, int32_t ncols, void* workspace,
                                                                      size_t workspace_bytes, vmcpde_stream stream) {
  VMC_REQUIRE(col0 >= 0 && ncols > 0, "vmcpde_eigh_cols: bad eigenvector range");
  return eigh_impl(S, n, ld, ev, VT, col0, ncols, workspace, workspace_bytes, stream);
}
