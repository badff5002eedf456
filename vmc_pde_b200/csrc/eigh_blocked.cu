// Blocked stages of the device eigensolver (vmcpde_eigh, replaces np.linalg.eigh at tdvp.py:61-64):
//
//  * tridiag_blocked: panel tridiagonalisation.  One persistent cooperative kernel per panel of nb columns
//    (one CTA per SM, two grid barriers per column).  Inside a panel the trailing matrix is only READ: the
//    matrix-vector product A v streams it once per column from HBM (8 B per element; the previous rank-2
//    formulation read and wrote it), the pending panel reflectors enter through skinny corrections, and the
//    rank-2nb trailing update is one FP64 tensor-core product (vmcpde_gemm_tn) per panel.
//    Matrix rows are dealt to CTAs in 4-row chunks, cyclically, so every SM keeps a fair share of the
//    shrinking trailing matrix; each CTA keeps the panel vectors of ITS rows in shared memory, so the only
//    data that crosses CTAs per column are 2i+2 partial dot products and the updated pivot row.
//  * backtransform_blocked: compact-WY application of the reflectors in blocks of 128,
//    V <- (I - Y T Y^T) V, as three tensor-core products per block.
//
// Row convention of eigh.cu: reflector v_j lives in row j of the work matrix, columns j+1.. (v_j[j+1] = 1).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

extern "C" int vmcpde_gemm_tn(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Out, int64_t ldo,
                              int32_t M, int32_t N, int64_t K, double alpha, double beta, vmcpde_stream stream);
extern "C" int vmcpde_gemm_tn_splitk(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Part, int64_t ldo,
                                     int32_t M, int32_t N, int64_t K, int32_t splits, vmcpde_stream stream);

namespace vmc {

namespace {

constexpr int kPT = 512;        // threads per CTA of the panel kernel
constexpr int kPW = kPT / 32;   // warps
constexpr int kRC = 4;          // rows per ownership chunk

__device__ __forceinline__ double wsum_(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum in a fixed order, result broadcast; sh >= 33 doubles
__device__ __forceinline__ double bsum_(double v, double* sh) {
  v = wsum_(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = (lane < nw) ? sh[lane] : 0.0;
  return wsum_(r);
}

// Barrier over all CTAs of the (cooperative) grid.  148 same-address atomics serialise at L2 (~27 cycles each, ~4000
// cycles per barrier on B200) and 148 x 148 flag polls hot-spot a handful of lines (measured: worse), so arrivals are
// spread over 16 counters on separate 128-byte lines and one warp per CTA polls their sum.  Arrival is a release
// reduction (covers the CTA's earlier writes through the preceding bar.sync); the poll is relaxed, followed by one
// acquire fence; the closing bar.sync extends the ordering to the whole CTA.  Counters only grow (epoch * grid).
constexpr int kBarCounters = 16, kBarStride = 32;
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr + (blockIdx.x % kBarCounters) * kBarStride) : "memory");
    const unsigned target = epoch * gridDim.x;
    unsigned sum;
    do {
      unsigned v = 0;
      if (threadIdx.x < kBarCounters)
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + threadIdx.x * kBarStride) : "memory");
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      sum = __shfl_sync(0xffffffffu, v, 0);
    } while (sum < target);
    __threadfence();
  }
  __syncthreads();
}

// 16-byte message {value.lo, tag, value.hi, tag}: only 8-byte accesses are single-copy atomic, so each half of the
// double travels with its own copy of the tag and the reader waits until both match (the line format of NCCL's LL
// protocol).  A 16-byte (value, tag) pair can tear: the tag may arrive before the value.
__device__ __forceinline__ void ll_store(double* slot, double v, unsigned tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"((unsigned)bits), "r"(tag),
               "r"((unsigned)(bits >> 32)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ double ll_load(const double* slot, unsigned tag) {
  unsigned lo, t1, hi, t2;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t1), "=r"(hi), "=r"(t2) : "l"(slot) : "memory");
  } while (t1 != tag || t2 != tag);
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

__device__ __forceinline__ int gcd16(int a) {  // gcd(a, 16) for a >= 1
  int g = 1;
  while (g < 16 && (a & g) == 0) g <<= 1;
  return g;
}

struct PanelArgs {
  double* A;
  int ld, n, j0, nb, nbp;  // panel columns j0 .. j0 + nbp - 1
  int j1;                  // first index kept in the masked copies X / Y (j0 + nb)
  int S;                   // owned slots per CTA (multiple of 4)
  double *d, *e, *tau;
  double* acol;            // [ld] updated pivot row
  double* part;            // [2 nb + 2][grid][2] partial dot products as (value, sequence) messages
  double* psig;            // [grid]
  double* vrows;           // [nb][nb + 1]: v_k(j0 + ii), ii in [0, nb]
  double* wrows;           // [nb][nb + 1]: w_k(j0 + ii)
  double* X;               // [2 nb][ld]: rows k: v_k, rows nb + k: w_k, zero for index < j1
  double* Y;               // [2 nb][ld]: rows k: w_k, rows nb + k: v_k, zero for index < j1
  double* Z;               // [grid][ld] column-part partial products of the symmetric mat-vec (NULL: full rows only)
  int sym_min_m;           // columns with a trailing size >= this read only the lower triangle
  int zs_global;           // the column-part accumulator of a CTA lives in its row of Z (large n: no room in shared memory)
  long long* prof;         // optional per-phase cycle counters of CTA 0 (VMCPDE_PANEL_PROFILE)
  unsigned* bar;           // barrier counters
  unsigned epoch0;         // barrier epochs completed before this launch
  double* totals;          // [2 nb + 1][2] reduced partials as (value, sequence) messages
};

#define VMC_PROF(k)                                                          \
  do {                                                                       \
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {                     \
      const long long t_ = clock64();                                        \
      a.prof[k] += t_ - tprof;                                               \
      tprof = t_;                                                            \
    }                                                                        \
  } while (0)

__global__ void __launch_bounds__(kPT, 1) tridiag_panel_kernel(const __grid_constant__ PanelArgs a) {
  long long tprof = clock64();
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, ld = a.ld, nb = a.nb, G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int S = a.S, ldo = nb + 1;
  const int nv = (n + 2) & ~1;  // vs[n] (and beyond) stay zero
  double* vs = sm;
  double* own_v = vs + nv;          // [S][ldo]
  double* own_w = own_v + S * ldo;  // [S][ldo]
  double* ys = own_w + S * ldo;     // [S]
  double* vown = ys + S;            // [S]
  double* PV = vown + S;            // [nb]
  double* PWs = PV + nb;            // [nb]
  double* rowV = PWs + nb;          // [nb]
  double* rowW = rowV + nb;         // [nb]
  double* ypart = rowW + nb;        // [S / 4][16][4]
  double* anext = ypart + S * 16;   // [S]
  double* red = anext + S;          // [64]
  // column-part accumulator of the symmetric mode: [nv] in shared memory, or this CTA's own row of Z when v already
  // fills shared memory (n > ~12k); every 64-column segment is touched by one warp only, so plain loads/stores suffice
  double* zs = a.zs_global ? a.Z + (size_t)b * ld : red + 64;

  auto row_of_slot = [&](int s) { return (b + (s >> 2) * G) * kRC + (s & 3); };
  const int nq = (n + kRC - 1) / kRC;                   // ownership chunks
  const int Lmy = nq > b ? (nq - b + G - 1) / G : 0;    // chunks of this CTA
  const int wseg = (n + G - 1) / G;                     // share of the replicated vector writes
  const int wlo = b * wseg, whi = min(n, wlo + wseg);

  for (int idx = tid; idx < 2 * S * ldo; idx += kPT) own_v[idx] = 0.0;
  for (int idx = tid; idx < nv; idx += kPT) vs[idx] = 0.0;
  unsigned target = a.epoch0;
  {  // prologue: pivot row of the first column (the trailing matrix is up to date)
    double s = 0.0;
    if (tid < S) {
      const int r = row_of_slot(tid);
      if (r >= a.j0 && r < n) {
        const double val = a.A[(size_t)a.j0 * ld + r];
        a.acol[r] = val;
        if (r >= a.j0 + 2) s = val * val;
      }
    }
    s = bsum_(s, red);
    if (tid == 0) a.psig[b] = s;
  }
  grid_sync(a.bar, target);

  for (int i = 0; i < a.nbp; ++i) {
    const int j = a.j0 + i, m = n - j - 1;
    const bool have = m >= 2;
    double tau_j = 0.0;
    // panel vectors at the next pivot row j + 1 (written one or more barriers ago)
    if (tid < i) {
      rowV[tid] = __ldcg(a.vrows + (size_t)tid * ldo + i + 1);
      rowW[tid] = __ldcg(a.wrows + (size_t)tid * ldo + i + 1);
    }
    // next pivot row of the (panel-start) trailing matrix on the rows of this CTA: issued now, used in phase B
    double anext_reg = 0.0;   // stays in flight during phase A; parked in shared memory before the first barrier
    if (tid < S) {
      const int r = row_of_slot(tid);
      if (i + 1 < a.nbp && r > j && r < n) anext_reg = __ldg(a.A + (size_t)(j + 1) * ld + r);
    }
    if (have) {
      // ---------------- phase A: reflector, y = A22 v on the rows of this CTA, partial dots ----------------
      const unsigned seq = (unsigned)j + 1u;   // message tag of this column
      // one round trip for everything the reflector needs: norm partials and the raw pivot row
      const double sg = tid < G ? __ldcg(a.psig + tid) : 0.0;
      for (int c = tid; c < n; c += kPT) vs[c] = c > j ? __ldcg(a.acol + c) : 0.0;
      const double sigma = bsum_(sg, red);
      VMC_PROF(0);
      const double alpha = vs[j + 1];
      double beta = alpha, scale = 0.0;
      if (sigma != 0.0) {
        beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
        tau_j = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      if (b == 0 && tid == 0) { a.d[j] = __ldcg(a.acol + j); a.e[j] = beta; a.tau[j] = tau_j; }
      __syncthreads();  // every thread has read alpha
      for (int c = j + 1 + tid; c < n; c += kPT) vs[c] = c == j + 1 ? 1.0 : vs[c] * scale;
      const bool sym = a.Z != nullptr && m >= a.sym_min_m;
      if (sym) for (int c = (j & ~1) + tid; c < (a.zs_global ? ((n + 1) & ~1) : nv); c += kPT) zs[c] = 0.0;
      __syncthreads();
      for (int c = wlo + tid; c < whi; c += kPT) {
        const double val = vs[c];
        if (c > j) a.A[(size_t)j * ld + c] = val;
        const double mv = c >= a.j1 ? val : 0.0;
        a.X[(size_t)i * ld + c] = mv;
        a.Y[(size_t)(nb + i) * ld + c] = mv;
      }
      if (b == 0 && tid <= nb) {
        const int c = a.j0 + tid;
        a.vrows[(size_t)i * ldo + tid] = c < n ? vs[c] : 0.0;
      }
      if (tid < S) {
        const int r = row_of_slot(tid);
        const double val = r < n ? vs[r] : 0.0;
        own_v[tid * ldo + i] = val;
        vown[tid] = val;
      }
      VMC_PROF(1);
      // y = A22 v: 4-row chunks x column segments over the warps
      const int qfirst = (j + 1) / kRC;
      const int lq0 = qfirst > b ? (qfirst - b + G - 1) / G : 0;
      const int kact = Lmy - lq0;
      const int cs0 = (j + 1) & ~1, ce0 = n & ~1;
      const int len = ce0 - cs0;
      int nseg = 1, segl = 64;
      if (sym) {
        // Lower triangle only: element (r, c), c < r, serves y(r) += A v(c) (row part, reduced inside the CTA) and
        // y(c) += A v(r) (column part, accumulated per CTA in zs and summed across CTAs after the barrier).
        // Chunk-outer / segment-inner: a warp owns the 64-column segments g = warp (mod 16), so zs needs no atomics.
        nseg = kPW;
        const int nsegs = (len + 63) >> 6;
        for (int cc = 0; cc < kact; ++cc) {
          const int ch = (j & 1) ? kact - 1 - cc : cc;   // odd columns walk backwards (L2 reuse)
          const int r0 = (b + (lq0 + ch) * G) * kRC;
          const int clast = min(r0 + 3, ce0 - 1);
          const int gmax = clast >= cs0 ? min(nsegs, ((clast - cs0) >> 6) + 1) : 0;
          const double* p0 = a.A + (size_t)min(r0, n - 1) * ld;
          const double* p1 = a.A + (size_t)min(r0 + 1, n - 1) * ld;
          const double* p2 = a.A + (size_t)min(r0 + 2, n - 1) * ld;
          const double* p3 = a.A + (size_t)min(r0 + 3, n - 1) * ld;
          const double vr0 = r0 < n ? vs[r0] : 0.0, vr1 = r0 + 1 < n ? vs[r0 + 1] : 0.0;
          const double vr2 = r0 + 2 < n ? vs[r0 + 2] : 0.0, vr3 = r0 + 3 < n ? vs[r0 + 3] : 0.0;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          auto masked = [&](int c, double2 av, double vr, int r, double2 va, double& sacc, double2& z) {
            if (c < r) { sacc = fma(av.x, va.x, sacc); z.x = fma(av.x, vr, z.x); }
            else if (c == r) sacc = fma(av.x, va.x, sacc);
            if (c + 1 < r) { sacc = fma(av.y, va.y, sacc); z.y = fma(av.y, vr, z.y); }
            else if (c + 1 == r) sacc = fma(av.y, va.y, sacc);
          };
          auto process = [&](int c, double2 a0, double2 a1, double2 a2, double2 a3) {
            const double2 va = *(const double2*)(vs + c);
            double2 z = *(double2*)(zs + c);
            if (c + 1 < r0) {  // strictly below the diagonal for all four rows
              s0 = fma(a0.x, va.x, s0); s0 = fma(a0.y, va.y, s0);
              s1 = fma(a1.x, va.x, s1); s1 = fma(a1.y, va.y, s1);
              s2 = fma(a2.x, va.x, s2); s2 = fma(a2.y, va.y, s2);
              s3 = fma(a3.x, va.x, s3); s3 = fma(a3.y, va.y, s3);
              z.x = fma(a0.x, vr0, z.x); z.x = fma(a1.x, vr1, z.x); z.x = fma(a2.x, vr2, z.x); z.x = fma(a3.x, vr3, z.x);
              z.y = fma(a0.y, vr0, z.y); z.y = fma(a1.y, vr1, z.y); z.y = fma(a2.y, vr2, z.y); z.y = fma(a3.y, vr3, z.y);
            } else {
              masked(c, a0, vr0, r0, va, s0, z); masked(c, a1, vr1, r0 + 1, va, s1, z);
              masked(c, a2, vr2, r0 + 2, va, s2, z); masked(c, a3, vr3, r0 + 3, va, s3, z);
            }
            *(double2*)(zs + c) = z;
          };
          int g = warp;
#pragma unroll 1
          for (; g + kPW < gmax; g += 2 * kPW) {  // two segments per trip: 8 x 16 B in flight per lane
            const int c = cs0 + (g << 6) + 2 * lane, c2 = c + (kPW << 6);
            const bool in2 = c2 < ce0;   // c < ce0 always holds here: a later segment exists
            const double2 a0 = __ldg((const double2*)(p0 + c)), a1 = __ldg((const double2*)(p1 + c));
            const double2 a2 = __ldg((const double2*)(p2 + c)), a3 = __ldg((const double2*)(p3 + c));
            double2 b0 = make_double2(0.0, 0.0), b1 = b0, b2 = b0, b3 = b0;
            if (in2) {
              b0 = __ldg((const double2*)(p0 + c2)); b1 = __ldg((const double2*)(p1 + c2));
              b2 = __ldg((const double2*)(p2 + c2)); b3 = __ldg((const double2*)(p3 + c2));
            }
            process(c, a0, a1, a2, a3);
            if (in2) process(c2, b0, b1, b2, b3);
          }
          if (g < gmax) {
            const int c = cs0 + (g << 6) + 2 * lane;
            if (c < ce0) {
              const double2 a0 = __ldg((const double2*)(p0 + c)), a1 = __ldg((const double2*)(p1 + c));
              const double2 a2 = __ldg((const double2*)(p2 + c)), a3 = __ldg((const double2*)(p3 + c));
              process(c, a0, a1, a2, a3);
            }
          }
          s0 = wsum_(s0); s1 = wsum_(s1); s2 = wsum_(s2); s3 = wsum_(s3);
          if (lane == 0) {
            double* yp = ypart + (ch * 16 + warp) * 4;
            yp[0] = s0; yp[1] = s1; yp[2] = s2; yp[3] = s3;
          }
        }
      } else {
      if (kact > 0 && len > 0) {
        nseg = 16 / gcd16(kact);
        while (nseg > 1 && len / nseg < 128) nseg >>= 1;
        segl = ((len + nseg - 1) / nseg + 63) & ~63;
      }
      const int items = kact > 0 ? kact * nseg : 0;
      // odd columns walk the items backwards: what the previous column read last is still in L2
      for (int it = warp; it < items; it += kPW) {
        const int item = (j & 1) ? items - 1 - it : it;
        const int ch = item % kact, sgi = item / kact;
        const int r0 = (b + (lq0 + ch) * G) * kRC;
        const double* p0 = a.A + (size_t)min(r0, n - 1) * ld;
        const double* p1 = a.A + (size_t)min(r0 + 1, n - 1) * ld;
        const double* p2 = a.A + (size_t)min(r0 + 2, n - 1) * ld;
        const double* p3 = a.A + (size_t)min(r0 + 3, n - 1) * ld;
        const int clo = cs0 + sgi * segl, chi = min(ce0, clo + segl);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = clo + 2 * lane;
#pragma unroll 1
        for (; c + 64 < chi; c += 128) {
          const double2 va = *(const double2*)(vs + c), vb = *(const double2*)(vs + c + 64);
          const double2 a0 = __ldg((const double2*)(p0 + c)), a1 = __ldg((const double2*)(p1 + c));
          const double2 a2 = __ldg((const double2*)(p2 + c)), a3 = __ldg((const double2*)(p3 + c));
          const double2 b0 = __ldg((const double2*)(p0 + c + 64)), b1 = __ldg((const double2*)(p1 + c + 64));
          const double2 b2 = __ldg((const double2*)(p2 + c + 64)), b3 = __ldg((const double2*)(p3 + c + 64));
          s0 = fma(a0.x, va.x, s0); s0 = fma(a0.y, va.y, s0); s0 = fma(b0.x, vb.x, s0); s0 = fma(b0.y, vb.y, s0);
          s1 = fma(a1.x, va.x, s1); s1 = fma(a1.y, va.y, s1); s1 = fma(b1.x, vb.x, s1); s1 = fma(b1.y, vb.y, s1);
          s2 = fma(a2.x, va.x, s2); s2 = fma(a2.y, va.y, s2); s2 = fma(b2.x, vb.x, s2); s2 = fma(b2.y, vb.y, s2);
          s3 = fma(a3.x, va.x, s3); s3 = fma(a3.y, va.y, s3); s3 = fma(b3.x, vb.x, s3); s3 = fma(b3.y, vb.y, s3);
        }
        if (c < chi) {
          const double2 va = *(const double2*)(vs + c);
          const double2 a0 = __ldg((const double2*)(p0 + c)), a1 = __ldg((const double2*)(p1 + c));
          const double2 a2 = __ldg((const double2*)(p2 + c)), a3 = __ldg((const double2*)(p3 + c));
          s0 = fma(a0.x, va.x, s0); s0 = fma(a0.y, va.y, s0);
          s1 = fma(a1.x, va.x, s1); s1 = fma(a1.y, va.y, s1);
          s2 = fma(a2.x, va.x, s2); s2 = fma(a2.y, va.y, s2);
          s3 = fma(a3.x, va.x, s3); s3 = fma(a3.y, va.y, s3);
        }
        if ((n & 1) && sgi == nseg - 1 && lane == 0) {  // odd last column
          const double vl = vs[n - 1];
          s0 = fma(p0[n - 1], vl, s0); s1 = fma(p1[n - 1], vl, s1); s2 = fma(p2[n - 1], vl, s2); s3 = fma(p3[n - 1], vl, s3);
        }
        s0 = wsum_(s0); s1 = wsum_(s1); s2 = wsum_(s2); s3 = wsum_(s3);
        if (lane == 0) {
          double* yp = ypart + (ch * 16 + sgi) * 4;
          yp[0] = s0; yp[1] = s1; yp[2] = s2; yp[3] = s3;
        }
      }
      }
      __syncthreads();
      VMC_PROF(2);
      double yvp = 0.0;   // this thread's share of y.v
      if (tid < S) {
        const int ch = (tid >> 2) - lq0, r = row_of_slot(tid);
        double s = 0.0;
        if (ch >= 0 && ch < kact && r > j && r < n) {
          for (int q = 0; q < nseg; ++q) s += ypart[(ch * 16 + q) * 4 + (tid & 3)];
          if (sym && (n & 1) && r == n - 1) s = fma(a.A[(size_t)r * ld + r], vs[r], s);  // diagonal outside the double2 range
        }
        ys[tid] = s;
        yvp = s * vown[tid];
      }
      if (sym) {  // publish the column part of this CTA
        double* zrow = a.Z + (size_t)b * ld;
        for (int c = j + 1 + tid; c < n; c += kPT) {
          const double z = zs[c];
          if (!a.zs_global) zrow[c] = z;   // global accumulator: already in place
          yvp = fma(z, vs[c], yvp);
        }
      }
      // partial dots with the panel vectors, y.v, and y at the next pivot row
      if (tid < 2 * i) {
        const double* arr = tid < i ? own_v : own_w;
        const int k = tid < i ? tid : tid - i;
        double s = 0.0;
        for (int q = 0; q < S; ++q) s = fma(arr[q * ldo + k], vown[q], s);
        ll_store(a.part + 2 * ((size_t)((tid < i ? 0 : nb) + k) * G + b), s, seq);
      }
      if (tid < S) anext[tid] = anext_reg;
      yvp = bsum_(yvp, red);
      if (tid == 0) ll_store(a.part + 2 * ((size_t)(2 * nb) * G + b), yvp, seq);
      if (tid < S && row_of_slot(tid) == j + 1) ll_store(a.part + 2 * (size_t)(2 * nb + 1) * G, ys[tid], seq);
      VMC_PROF(3);
      // The partial dots travel as (value, sequence) messages, so they need no barrier; the bulk column parts of the
      // symmetric mat-vec do.
      if (sym) grid_sync(a.bar, target);
      VMC_PROF(4);
      // ---------------- phase B: reduce the partials ----------------
      if (sym) {
        // y on the rows of this CTA += column parts of all CTAs (8 lanes per row, fixed order)
        for (int base = 0; base < S; base += kPT / 8) {
          const int sl = base + (tid >> 3), kp = tid & 7;
          double zsum = 0.0;
          const int r = sl < S ? row_of_slot(sl) : n;
          if (r > j && r < n) {
            // up to 160 CTAs: 20 loads per lane, straight-line so that they are all in flight together
            double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
            const double* zc = a.Z + r;
#pragma unroll
            for (int u = 0; u < 20; u += 4) {
              const int b0 = kp + 8 * u;
              if (b0 < G) z0 += __ldcg(zc + (size_t)b0 * ld);
              if (b0 + 8 < G) z1 += __ldcg(zc + (size_t)(b0 + 8) * ld);
              if (b0 + 16 < G) z2 += __ldcg(zc + (size_t)(b0 + 16) * ld);
              if (b0 + 24 < G) z3 += __ldcg(zc + (size_t)(b0 + 24) * ld);
            }
            for (int bb = kp + 160; bb < G; bb += 8) z0 += __ldcg(zc + (size_t)bb * ld);
            zsum = (z0 + z1) + (z2 + z3);
          }
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
          if (kp == 0 && sl < S) ys[sl] += zsum;
        }
        if (warp == kPW - 2) {  // column parts of the next pivot row
          double zsum = 0.0;
          for (int bb = lane; bb < G; bb += 32) zsum += __ldcg(a.Z + (size_t)bb * ld + j + 1);
          zsum = wsum_(zsum);
          if (lane == 0) red[42] = zsum;
        }
      } else if (tid == 0) {
        red[42] = 0.0;
      }
      {  // Value t of the 2 i + 1 reduced partials is summed by CTA t (warp 0, coalesced, fixed order) and published as
         // a (value, sequence) message; every CTA then collects all of them.  Two L2 round trips, no all-to-all reads.
        const int nred = 2 * i + 1;
        for (int t = b; t < nred; t += G) {
          if (warp == 0) {
            const int kk = t < i ? t : (t < 2 * i ? nb + (t - i) : 2 * nb);
            const double* src = a.part + 2 * (size_t)kk * G;
            double sacc = 0.0;
            for (int q = lane; q < G; q += 32) sacc += ll_load(src + 2 * q, seq);   // waits for CTA q's message
            sacc = wsum_(sacc);
            if (lane == 0) ll_store(a.totals + 2 * t, sacc, seq);
          }
        }
        VMC_PROF(9);
        if (tid == kPT - 1) red[41] = ll_load(a.part + 2 * (size_t)(2 * nb + 1) * G, seq);   // y at the next pivot row
        for (int t = tid; t < nred; t += kPT) {
          const double v = ll_load(a.totals + 2 * t, seq);
          if (t < i) PV[t] = v;
          else if (t < 2 * i) PWs[t - i] = v;
          else red[40] = v;
        }
        VMC_PROF(10);
      }
    } else {
      // no reflector for the last two columns: d / e only, zero panel vectors
      if (b == 0 && tid == 0) {
        a.d[j] = __ldcg(a.acol + j);
        a.e[j] = m == 1 ? __ldcg(a.acol + j + 1) : 0.0;
        a.tau[j] = 0.0;
      }
      for (int c = wlo + tid; c < whi; c += kPT) {
        a.X[(size_t)i * ld + c] = 0.0; a.Y[(size_t)(nb + i) * ld + c] = 0.0;
        a.X[(size_t)(nb + i) * ld + c] = 0.0; a.Y[(size_t)i * ld + c] = 0.0;
      }
      if (b == 0 && tid <= nb) { a.vrows[(size_t)i * ldo + tid] = 0.0; a.wrows[(size_t)i * ldo + tid] = 0.0; }
      if (tid < S) { vown[tid] = 0.0; ys[tid] = 0.0; anext[tid] = anext_reg; }
      if (tid == 0) { red[40] = 0.0; red[41] = 0.0; red[42] = 0.0; }
    }
    __syncthreads();
    VMC_PROF(5);
    // scalars of the column: one warp reduces, everybody reads them back (16 warps repeating the loop would
    // queue on the FP64 pipe)
    if (warp == 0) {
      double pd = 0.0, cd = 0.0;
      for (int k = lane; k < i; k += 32) {
        pd = fma(PV[k], PWs[k], pd);
        cd = fma(rowV[k], PWs[k], cd);
        cd = fma(rowW[k], PV[k], cd);
      }
      pd = wsum_(pd); cd = wsum_(cd);
      if (lane == 0) { red[43] = pd; red[44] = cd; }
    }
    __syncthreads();
    const double pvdot = red[43], c1 = red[44];
    const double pv = have ? tau_j * (red[40] - 2.0 * pvdot) : 0.0;   // p.v with p = tau * (corrected y)
    const double w1 = have ? tau_j * (red[41] + red[42] - c1) - 0.5 * tau_j * pv : 0.0;  // w at the next pivot row
    VMC_PROF(6);
    // w on the rows of this CTA and the next pivot row, 8 lanes per slot
    double sg2 = 0.0;
    for (int base = 0; base < S; base += kPT / 8) {
      const int s = base + (tid >> 3), kp = tid & 7;
      double accw = 0.0, acca = 0.0;
      if (s < S) {
        for (int k = kp; k < i; k += 8) {
          const double ov = own_v[s * ldo + k], ow = own_w[s * ldo + k];
          if (have) { accw = fma(ov, PWs[k], accw); accw = fma(ow, PV[k], accw); }
          acca = fma(rowV[k], ow, acca); acca = fma(rowW[k], ov, acca);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        accw += __shfl_xor_sync(0xffffffffu, accw, o);
        acca += __shfl_xor_sync(0xffffffffu, acca, o);
      }
      if (kp == 0 && s < S) {
        const int r = row_of_slot(s);
        const double vv = vown[s];
        double w = 0.0;
        if (have && r > j && r < n) w = tau_j * (ys[s] - accw) - 0.5 * tau_j * pv * vv;
        own_w[s * ldo + i] = w;
        if (r < n && have) {
          const double mw = r >= a.j1 ? w : 0.0;
          a.X[(size_t)(nb + i) * ld + r] = mw;
          a.Y[(size_t)i * ld + r] = mw;
          if (r >= a.j0 && r <= a.j0 + nb) a.wrows[(size_t)i * ldo + r - a.j0] = w;
        }
        if (i + 1 < a.nbp && r > j && r < n) {  // pivot row j + 1 with all panel reflectors applied
          const double an = anext[s] - acca - (w + w1 * vv);
          a.acol[r] = an;
          if (r >= j + 3) sg2 += an * an;
        }
      }
    }
    sg2 = bsum_(sg2, red);
    if (tid == 0) a.psig[b] = sg2;
    VMC_PROF(7);
    grid_sync(a.bar, target);
    VMC_PROF(8);
  }
}

// A[j][c] (c > j, j <= n-3) holds v_j: AT[c][j] = v_j(c), everything else zero; A is masked in place the same way.
__global__ void __launch_bounds__(256) mask_transpose_kernel(double* __restrict__ A, double* __restrict__ AT, int ld, int n) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8) {
    const int j = r0 + r, c = c0 + tx;
    const double raw = A[(size_t)j * ld + c];
    const double v = (c > j && j + 2 < n && c < n) ? raw : 0.0;
    A[(size_t)j * ld + c] = v;
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) AT[(size_t)(c0 + r) * ld + r0 + tx] = tile[tx][r];
}

__global__ void __launch_bounds__(256) transpose_sq_kernel(const double* __restrict__ A, double* __restrict__ B, int ld) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = A[(size_t)(r0 + r) * ld + c0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) B[(size_t)(c0 + r) * ld + r0 + tx] = tile[tx][r];
}

// B[c0 + c][r] = A[r][c0 + c] for r < rows, c < cols (multiples of 32): columns [c0, c0 + cols) of A become rows of B
__global__ void __launch_bounds__(256) transpose_cols_kernel(const double* __restrict__ A, double* __restrict__ B, int ld, int c0) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, cc = c0 + blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = A[(size_t)(r0 + r) * ld + cc + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) B[(size_t)(cc + r) * ld + r0 + tx] = tile[tx][r];
}

// G[r][c] = sum_s Part[s][r][c] for the 128 x ncols split-K slices of Y^T V (fixed order)
__global__ void __launch_bounds__(256) sum_slices_kernel(const double* __restrict__ part, int splits, int ld, int ncols,
                                                         double* __restrict__ G) {
  const int r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += gridDim.x * blockDim.x) {
    double sacc = 0.0;
    for (int q = 0; q < splits; ++q) sacc += part[((size_t)q * 128 + r) * ld + c];
    G[(size_t)r * ld + c] = sacc;
  }
}

// Partial Gram of one 128-reflector block over a slice of 1024 coordinates:
// part[blk][slice][k][k'] = sum_{c in slice} AT[c][128 blk + k] AT[c][128 blk + k']   (DMMA, 8 warps x 64x32)
constexpr int kBK = 128;       // reflectors per back-transform block
constexpr int kSlice = 1024;   // coordinates per partial Gram
constexpr int kYld = 132;      // padded shared-memory stride (conflict-free fragment loads)
__device__ __forceinline__ void dmma_b(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) block_gram_kernel(const double* __restrict__ AT, int ld, int np, double* __restrict__ part,
                                                         int max_slices) {
  __shared__ double Ys[16][kYld];
  const int blk = blockIdx.y, sl = blockIdx.x;
  const int cbeg = blk * kBK + sl * kSlice;
  if (cbeg >= np) return;
  const int cend = min(np, cbeg + kSlice);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const double* src = AT + (size_t)blk * kBK;
  for (int c0 = cbeg; c0 < cend; c0 += 16) {
    for (int idx = threadIdx.x; idx < 16 * kBK; idx += 256) {
      const int r = idx >> 7, k = idx & 127;
      Ys[r][k] = src[(size_t)(c0 + r) * ld + k];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      double av[8], bv[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = Ys[s * 4 + t][wm * 64 + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Ys[s * 4 + t][wn * 32 + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_b(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
    }
    __syncthreads();
  }
  double* out = part + ((size_t)blk * max_slices + sl) * kBK * kBK;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = wm * 64 + i * 8 + g, c = wn * 32 + j * 8 + 2 * t;
      *(double2*)(out + (size_t)r * kBK + c) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// T of the compact-WY form H_a ... H_{a+127} = I - Y T Y^T (LAPACK dlarft, forward / columnwise), one CTA per block.
// Tt[blk][k'][k] = T[k][k'] (transposed, the layout vmcpde_gemm_tn wants for T * G1).
__global__ void __launch_bounds__(128) larft_kernel(const double* __restrict__ part, int max_slices, int np,
                                                    const double* __restrict__ tau, int n, double* __restrict__ Tt) {
  extern __shared__ double Gm[];  // [128][129]; the upper triangle turns into T column by column
  __shared__ double colv[kBK];
  const int blk = blockIdx.x, tid = threadIdx.x;
  const int nsl = min(max_slices, (np - blk * kBK + kSlice - 1) / kSlice);
  for (int idx = tid; idx < kBK * kBK; idx += 128) {
    double s = 0.0;
    for (int q = 0; q < nsl; ++q) s += part[((size_t)blk * max_slices + q) * kBK * kBK + idx];
    Gm[(idx >> 7) * (kBK + 1) + (idx & 127)] = s;
  }
  __syncthreads();
  for (int i = 0; i < kBK; ++i) {
    const int j = blk * kBK + i;
    const double t = j < n ? tau[j] : 0.0;
    // colv[r] = -t * sum_{l = r}^{i-1} T[r][l] * Gm[l][i]
    double s = 0.0;
    if (tid < i) {
      for (int l = tid; l < i; ++l) s = fma(Gm[tid * (kBK + 1) + l], Gm[l * (kBK + 1) + i], s);
      colv[tid] = -t * s;
    }
    __syncthreads();
    if (tid < i) Gm[tid * (kBK + 1) + i] = colv[tid];
    if (tid == i) Gm[i * (kBK + 1) + i] = t;
    __syncthreads();
  }
  double* out = Tt + (size_t)blk * kBK * kBK;
  for (int idx = tid; idx < kBK * kBK; idx += 128) {
    const int kp = idx >> 7, k = idx & 127;  // out[k'][k] = T[k][k'] (zero below the diagonal of T)
    out[idx] = k <= kp ? Gm[k * (kBK + 1) + kp] : 0.0;
  }
}

size_t al256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

// --------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------
struct BlockedPlan {
  int nb = 0, S = 0, grid = 0;
  size_t smem = 0;
  bool sym = false;        // symmetric (lower-triangle) mat-vec available
  bool zs_global = false;  // ... with its column-part accumulator in global memory
};

static bool plan_panel(int n, BlockedPlan* p) {
  int dev = 0, coop = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (!coop) return false;
  const int G = num_sms();
  const int nq = (n + kRC - 1) / kRC;
  const int S = ((nq + G - 1) / G) * kRC;
  const int nv = (n + 2) & ~1;
  for (int nb = 64; nb >= 16; nb >>= 1) {
    const size_t doubles = (size_t)nv + 2 * (size_t)S * (nb + 1) + 3 * S + 4 * nb + (size_t)S * 16 + 64;
    if (doubles * 8 <= (size_t)max_smem) {
      p->nb = nb; p->S = S; p->grid = G; p->smem = doubles * 8;
      if (!getenv("VMCPDE_EIGH_NOSYM")) {
        p->sym = true;
        if (nb == 64 && (doubles + nv) * 8 <= (size_t)max_smem && !getenv("VMCPDE_EIGH_ZS_GLOBAL")) p->smem = (doubles + nv) * 8;
        else p->zs_global = true;
      }
      return true;
    }
  }
  return false;
}

bool blocked_eigh_supported(int n, int ld) {
  if (getenv("VMCPDE_EIGH_LEGACY")) return false;
  if (n < 384 || ld % 128 != 0 || ld < (n + 127) / 128 * 128) return false;
  BlockedPlan p;
  return plan_panel(n, &p);
}

size_t blocked_tridiag_scratch_bytes(int n, int ld) {
  BlockedPlan p;
  if (!plan_panel(n, &p)) return 0;
  const int nb = p.nb, G = p.grid;
  return al256((size_t)ld * 8) + al256((size_t)(2 * nb + 2) * G * 16) + al256((size_t)G * 8) +
         2 * al256((size_t)nb * (nb + 1) * 8) + 2 * al256((size_t)2 * nb * ld * 8) + al256((size_t)kBarCounters * kBarStride * 4) +
         al256((size_t)(2 * nb + 1) * 16) +
         (p.sym ? al256((size_t)G * ld * 8) : 0);
}

// launches issued by the two blocked stages (reported through vmcpde_eigh_launch_count)
int blocked_tridiag_launches(int n) {
  BlockedPlan p;
  if (!plan_panel(n, &p)) return 0;
  const int panels = (n + p.nb - 1) / p.nb;
  return 2 * panels - 1;
}
int blocked_backtransform_launches(int n) { return 5 + 4 * ((n + kBK - 1) / kBK); }   // upper bound (split-K adds one per block)

// A: np x ld work matrix (np = n rounded up to 128 rows must be allocated), destroyed: on exit row j holds v_j.
int tridiag_blocked(double* A, int n, int ld, double* d, double* e, double* tau, void* scratch, size_t scratch_bytes,
                    cudaStream_t s) {
  BlockedPlan p;
  if (!plan_panel(n, &p)) return set_error(VMCPDE_EUNSUPPORTED, "tridiag_blocked: no cooperative launch / shared memory");
  if (scratch_bytes < blocked_tridiag_scratch_bytes(n, ld)) return set_error(VMCPDE_EINVAL, "tridiag_blocked: scratch too small");
  const int nb = p.nb, G = p.grid, np = (n + 127) / 128 * 128;
  uint8_t* wp = (uint8_t*)scratch;
  auto take = [&](size_t bytes) { void* q = wp; wp += al256(bytes); return q; };
  PanelArgs a{};
  a.A = A; a.ld = ld; a.n = n; a.nb = nb; a.S = p.S; a.d = d; a.e = e; a.tau = tau;
  a.acol = (double*)take((size_t)ld * 8);
  a.part = (double*)take((size_t)(2 * nb + 2) * G * 16);
  a.psig = (double*)take((size_t)G * 8);
  a.vrows = (double*)take((size_t)nb * (nb + 1) * 8);
  a.wrows = (double*)take((size_t)nb * (nb + 1) * 8);
  a.X = (double*)take((size_t)2 * nb * ld * 8);
  a.Y = (double*)take((size_t)2 * nb * ld * 8);
  const int panels = (n + nb - 1) / nb;
  unsigned* bars = (unsigned*)take((size_t)kBarCounters * kBarStride * 4);
  a.totals = (double*)take((size_t)(2 * nb + 1) * 16);
  long long* prof = nullptr;
  if (getenv("VMCPDE_PANEL_PROFILE")) {
    VMC_CUDA_CHECK(cudaMalloc(&prof, 16 * sizeof(long long)));
    VMC_CUDA_CHECK(cudaMemsetAsync(prof, 0, 16 * sizeof(long long), s));
  }
  a.prof = prof;
  VMC_CUDA_CHECK(cudaMemsetAsync(a.acol, 0, (size_t)(wp - (uint8_t*)a.acol), s));
  a.Z = p.sym ? (double*)take((size_t)G * ld * 8) : nullptr;
  a.zs_global = p.zs_global ? 1 : 0;
  a.sym_min_m = 1536;
  if (const char* e_ = getenv("VMCPDE_EIGH_SYM_MIN_M")) a.sym_min_m = atoi(e_);
  VMC_CUDA_CHECK(cudaFuncSetAttribute(tridiag_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  unsigned epoch = 0;
  for (int pi = 0; pi < panels; ++pi) {
    a.j0 = pi * nb;
    a.nbp = min(nb, n - a.j0);
    a.j1 = a.j0 + nb;
    a.bar = bars;
    a.epoch0 = epoch;
    epoch += 1;   // barriers of this launch: prologue, one per column, one more per column in symmetric mode
    for (int i = 0; i < a.nbp; ++i) {
      const int m = n - (a.j0 + i) - 1;
      epoch += (m >= 2 && a.Z != nullptr && m >= a.sym_min_m) ? 2 : 1;
    }
    void* kargs[] = {(void*)&a};
    VMC_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)tridiag_panel_kernel, dim3(G), dim3(kPT), kargs, p.smem, s));
    if (a.j1 < n) {
      const int ja = a.j1 & ~127;
      if (int rc = vmcpde_gemm_tn(a.X + ja, ld, a.Y + ja, ld, A + (size_t)ja * ld + ja, ld, np - ja, np - ja, 2 * nb, -1.0, 1.0,
                                  (vmcpde_stream)s))
        return rc;
    }
  }
  VMC_LAUNCH_CHECK("tridiag_blocked");
  if (prof) {  // debugging aid: synchronises
    long long h[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    const char* names[11] = {"A: loads + norm", "A: build v, writes", "A: mat-vec", "A: y, partial dots", "barrier 1",
                             "B: rest of gather", "B: scalars", "B: w, next pivot row", "barrier 2",
                             "B: sym gather + owner reduce", "B: collect totals"};
    long long tot = 0;
    for (int k = 0; k < 11; ++k) tot += h[k];
    fprintf(stderr, "[tridiag_panel_kernel n=%d] CTA 0 cycles per column by phase (total %.0f):\n", n, (double)tot / n);
    for (int k = 0; k < 11; ++k) fprintf(stderr, "   %-30s %9.0f  %5.1f%%\n", names[k], (double)h[k] / n, 100.0 * h[k] / tot);
  }
  return 0;
}

size_t blocked_backtransform_scratch_bytes(int n, int ld) {
  const int np = (n + 127) / 128 * 128, nblk = np / kBK;
  const int max_slices = (np + kSlice - 1) / kSlice;
  return al256((size_t)nblk * max_slices * kBK * kBK * 8) + al256((size_t)nblk * kBK * kBK * 8) + 2 * al256((size_t)kBK * ld * 8) +
         al256((size_t)16 * kBK * ld * 8);   // split-K slices of Y^T V
}

// V = H_0 ... H_{n-3} Z.  ZT (np x ld, rows = eigenvectors of the tridiagonal, zero padded) is consumed;
// A holds the reflectors (row convention) and is masked in place; AT and Zn are np x ld work matrices;
// VT (rows = eigenvectors) may alias ZT or AT.  Only the eigenvectors [col0, col0 + ncols) (multiples of 128) are
// transformed and written (rows col0.. of VT): ranks of a multi-GPU solve each take a slice.
int backtransform_blocked(double* A, const double* tau, int n, int ld, const double* ZT, double* Zn, double* AT,
                          void* scratch, size_t scratch_bytes, double* VT, int col0, int ncols, cudaStream_t s) {
  const int np = (n + 127) / 128 * 128, nblk = np / kBK;
  const int max_slices = (np + kSlice - 1) / kSlice;
  if (scratch_bytes < blocked_backtransform_scratch_bytes(n, ld)) return set_error(VMCPDE_EINVAL, "backtransform_blocked: scratch too small");
  if (col0 < 0 || ncols <= 0 || col0 % 128 || ncols % 128 || col0 + ncols > np)
    return set_error(VMCPDE_EINVAL, "backtransform_blocked: eigenvector range must be 128-aligned inside the padded size");
  const dim3 tgrid(np / 32, np / 32);
  transpose_sq_kernel<<<tgrid, 256, 0, s>>>(ZT, Zn, ld);
  mask_transpose_kernel<<<tgrid, 256, 0, s>>>(A, AT, ld, n);
  uint8_t* wp = (uint8_t*)scratch;
  auto take = [&](size_t bytes) { void* q = wp; wp += al256(bytes); return q; };
  double* part = (double*)take((size_t)nblk * max_slices * kBK * kBK * 8);
  double* Tt = (double*)take((size_t)nblk * kBK * kBK * 8);
  double* G1 = (double*)take((size_t)kBK * ld * 8);
  double* G2 = (double*)take((size_t)kBK * ld * 8);
  double* Gs = (double*)take((size_t)16 * kBK * ld * 8);
  // Y^T V has only ncols / 128 output tiles: split its contraction so that tiles x splits fills the SMs
  const int g1_tiles = ncols / 128;
  int splits = g1_tiles >= num_sms() / 2 ? 1 : min(16, num_sms() / g1_tiles);
  if (getenv("VMCPDE_BT_NOSPLIT")) splits = 1;
  block_gram_kernel<<<dim3(max_slices, nblk), 256, 0, s>>>(AT, ld, np, part, max_slices);
  const size_t lsm = (size_t)kBK * (kBK + 1) * 8;
  VMC_CUDA_CHECK(cudaFuncSetAttribute(larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
  larft_kernel<<<nblk, 128, lsm, s>>>(part, max_slices, np, tau, n, Tt);
  VMC_LAUNCH_CHECK("backtransform_blocked setup");
  double* Zc = Zn + col0;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int c0 = blk * kBK, K = np - c0;
    // G1 = Y^T Zn  (128 x ncols), contraction over the coordinates c >= c0
    const int sp = min(splits, max(1, K / 256));   // at least 16 pipeline steps per slice
    if (sp > 1) {
      if (int rc = vmcpde_gemm_tn_splitk(AT + (size_t)c0 * ld + c0, ld, Zc + (size_t)c0 * ld, ld, Gs, ld, kBK, ncols, K, sp, (vmcpde_stream)s)) return rc;
      sum_slices_kernel<<<dim3(max(1, min(8, ncols / 256)), kBK), 256, 0, s>>>(Gs, sp, ld, ncols, G1);
    } else if (int rc = vmcpde_gemm_tn(AT + (size_t)c0 * ld + c0, ld, Zc + (size_t)c0 * ld, ld, G1, ld, kBK, ncols, K, 1.0, 0.0, (vmcpde_stream)s)) return rc;
    // G2 = T G1
    if (int rc = vmcpde_gemm_tn(Tt + (size_t)blk * kBK * kBK, kBK, G1, ld, G2, ld, kBK, ncols, kBK, 1.0, 0.0, (vmcpde_stream)s)) return rc;
    // Zn[c0:, range] -= Y G2
    if (int rc = vmcpde_gemm_tn(A + (size_t)c0 * ld + c0, ld, G2, ld, Zc + (size_t)c0 * ld, ld, K, ncols, kBK, -1.0, 1.0, (vmcpde_stream)s)) return rc;
  }
  transpose_cols_kernel<<<dim3(ncols / 32, np / 32), 256, 0, s>>>(Zn, VT, ld, col0);
  VMC_LAUNCH_CHECK("backtransform_blocked");
  return 0;
}

}  // namespace vmc
