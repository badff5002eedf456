// Weighted Gram build  S_m += O^T diag(w_m) O  on the FP64 tensor path (DMMA.8x8x4) of sm_100a.
//
// Replaces mpi_wrapper._cov_helper_without_p / global_covariance (mpi_wrapper.py:21-25,248-274) as used by
// tdvp.py:46 (S0), :47 (SExp, weights logp^2) and the EOdata@V covariance of tdvp.py:68-70 (weights dE^2).
//
// Design (B200-first, not a translation of the XLA dot):
//  * tcgen05/UMMA has no FP64 kind, so the tensor path for FP64 on Blackwell is mma.sync DMMA; the kernel
//    is built around it: 8 consumer warps, each a 64x32 register tile (64 FP64 accumulators / thread).
//  * O is row-major [samples][Pp] (the contraction index is the slow one).  One TMA 3-D box
//    {16 doubles, KC samples, 8 column groups} per operand and stage lands a 128-column x KC-sample
//    panel in shared memory as [column group][sample][16 doubles] with the 128-byte swizzle; mapping the
//    DMMA k index to samples {0,2,4,6}/{1,3,5,7} of each 8-sample group makes every fragment load
//    bank-conflict free without padding.  Both operands of a Gram come from the same matrix, so
//    diagonal tiles load a single panel.
//  * Persistent CTAs (one per SM) walk the upper-triangular tile pairs of all requested matrices,
//    ordered so that concurrently running CTAs share column panels in L2.  A producer warp runs the
//    TMA pipeline (mbarrier full/empty ring); consumers never touch global memory until the epilogue.
//  * Row weights (and nothing else) are applied to the B fragment in registers, so S0, SExp and the SNR
//    covariance read the same centred O with no N x P temporaries (the reference materialises three).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "common.cuh"

namespace vmc {

constexpr int kBM = 128;            // tile rows (parameters a)
constexpr int kBN = 128;            // tile cols (parameters b)
constexpr int kKC = 16;             // samples per pipeline stage
constexpr int kStages = 6;
constexpr int kConsumerWarps = 8;
constexpr int kGramThreads = (kConsumerWarps + 4) * 32;  // 2 consumer warpgroups + 1 producer warpgroup
constexpr int kPanelBytes = kBM * kKC * 8;  // 16 KB
constexpr int kWBytes = kKC * 8;            // 128 B of row weights
constexpr int kStageBytes = 2 * kPanelBytes + 1024;  // A, B, weights (padded to keep 1 KB alignment)
constexpr int kMaxMats = 16;   // weighted matrices of one SYRK launch (<= 4 are used) or K-slices of a split-K product

struct GramArgs {
  double* S[kMaxMats];
  const double* w[kMaxMats];
  int n_mats;
  int tiles;      // SYRK: Pp / 128 ; GEMM: tile rows (M / 128)
  int tiles_n;    // GEMM: tile cols (N / 128)
  int Pp;         // leading dimension of the outputs
  int full;       // 0: upper-triangular tile pairs of X^T X ; 1: all tiles of X^T Y (second tensor map)
  int super;      // supertile edge of the SYRK enumeration (tiles)
  int ksplit;     // full mode only: the n_mats "matrices" are K-slices of one product, slice m -> its own output S[m]
  int tail_ks;    // SYRK mode: the n_items % grid items of the last, partly filled round are each split into tail_ks K-slices
                  // (0: off) so that the round costs 1 / tail_ks of an item; slices add their result in slice order
  int wave_sync;  // 1: producers rendezvous at every work item (keeps the CTAs of a wave inside one L2 window)
  double alpha, beta;  // out = alpha * acc + beta * out
  long long n;    // contraction length (samples)
};

// Arrival counter of the per-wave rendezvous of the SYRK producers (performance only, see gram_kernel); zeroed before
// every launch that uses it.  One per device; launches that use it must not overlap on one device.
__device__ unsigned g_wave_counter;
// Turn counters of the split tail items (see gram_kernel): g_tail_turn[t] = number of K-slices of tail item t already added.
__device__ unsigned g_tail_turn[256];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// work item -> (matrix, tile row ti <= tile col tj).  Matrices innermost, so the weighted variants of one tile pair run
// side by side and share both panels in L2; tile pairs are enumerated supertile by supertile (G x G tile blocks, G chosen
// so that one supertile x n_mats ~ one wave of CTAs): the CTAs of a wave then stream 2 G column panels instead of
// G^2 + 1, which keeps the panel window of a wave inside the 126 MB L2 even when the CTAs drift apart in the sample index.
__device__ __forceinline__ void decode_item(long long item, const GramArgs& a, int& mat, int& ti, int& tj) {
  const int n_mats = a.n_mats, tiles = a.tiles;
  mat = (int)(item % n_mats);
  long long p = item / n_mats;
  if (a.full) { ti = (int)(p / a.tiles_n); tj = (int)(p % a.tiles_n); return; }
  const int G = a.super;
  const int ST = (tiles + G - 1) / G;
  for (int SI = 0; SI < ST; ++SI) {
    const int r0 = SI * G, rn = min(G, tiles - r0);
    long long cnt = (long long)rn * (rn + 1) / 2;   // diagonal supertile: its upper triangle, row by row
    if (p < cnt) {
      int r = 0;
      while (p >= (long long)(rn - r)) { p -= (rn - r); ++r; }
      ti = r0 + r;
      tj = r0 + r + (int)p;
      return;
    }
    p -= cnt;
    for (int SJ = SI + 1; SJ < ST; ++SJ) {
      const int c0 = SJ * G, cn = min(G, tiles - c0);
      cnt = (long long)rn * cn;
      if (p < cnt) { ti = r0 + (int)(p / cn); tj = c0 + (int)(p % cn); return; }
      p -= cnt;
    }
  }
  ti = tj = 0;  // not reached: item < n_items
}

__global__ void __launch_bounds__(kGramThreads, 1)
gram_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapB,
            const __grid_constant__ GramArgs args) {
  // dynamic shared memory is the only shared allocation, so it starts 1 KB aligned (128B-swizzle atom)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = (uint64_t*)(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long n_items = args.full ? (long long)args.n_mats * args.tiles * args.tiles_n
                                      : (long long)args.n_mats * args.tiles * (args.tiles + 1) / 2;
  const int k_iters = (int)(args.n / kKC);
  // Work units of this CTA: its items of the full rounds, then (tail_ks > 0) one K-slice of an item of the last round.
  const long long main_items = args.tail_ks ? n_items - n_items % gridDim.x : n_items;
  const int tail_items = (int)(n_items - main_items);
  auto get_unit = [&](long long u, long long& item, int& kb, int& ke) -> bool {   // tail unit <=> item >= main_items
    item = (long long)blockIdx.x + u * gridDim.x;
    kb = 0; ke = k_iters;
    if (item < main_items) return true;
    const long long u_main = main_items > (long long)blockIdx.x ? (main_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (args.tail_ks && u == u_main && (int)blockIdx.x < tail_items * args.tail_ks) {
      const int tidx = (int)blockIdx.x / args.tail_ks, tslice = (int)blockIdx.x % args.tail_ks;
      item = main_items + tidx;
      const int kper = (k_iters + args.tail_ks - 1) / args.tail_ks;
      kb = tslice * kper; ke = min(k_iters, kb + kper);
      return true;
    }
    return false;
  };

  if (warp >= kConsumerWarps) {
    // ===== producer warpgroup: hands its registers to the consumers; one elected lane drives TMA =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kConsumerWarps && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned wave_target = 0;
      for (long long u = 0;; ++u) {
        long long item; int k_begin, k_end;
        if (!get_unit(u, item, k_begin, k_end)) break;
        if (args.wave_sync && item < main_items) {
          // All CTAs take items of equal length, but they drift apart over a 1.5 s launch and then miss each other's panels
          // in L2.  The producers therefore meet before every item.  This is an optimisation only: the wait is bounded,
          // so nothing depends on the CTAs being co-resident.
          const long long wave_first = item - blockIdx.x;
          const long long left = n_items - wave_first;
          wave_target += (unsigned)(left < (long long)gridDim.x ? left : (long long)gridDim.x);
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(&g_wave_counter) : "memory");
          const long long t0 = clock64();
          unsigned seen;
          do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&g_wave_counter) : "memory");
          } while (seen < wave_target && clock64() - t0 < 4000000ll);   // <= ~2 ms
        }
        int mat, ti, tj;
        decode_item(item, args, mat, ti, tj);
        const bool diag = !args.full && (ti == tj);
        const double* w = args.w[mat];
        const uint32_t bytes = kPanelBytes + (diag ? 0 : kPanelBytes) + (w ? kWBytes : 0);
        if (args.ksplit) {
          const int kper = (k_iters + args.n_mats - 1) / args.n_mats;
          k_begin = mat * kper; k_end = min(k_iters, k_begin + kper);
        }
        for (int k = k_begin; k < k_end; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          mbar_expect_tx(&full[stage], bytes);
          tma_load_3d(st, &tmap, &full[stage], 0, k * kKC, ti * (kBM / 16));
          if (!diag) tma_load_3d(st + kPanelBytes, args.full ? &tmapB : &tmap, &full[stage], 0, k * kKC, tj * (kBN / 16));
          if (w) bulk_load_1d(st + 2 * kPanelBytes, w + (long long)k * kKC, kWBytes, &full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumers: 2 (M) x 4 (N) warps, warp tile 64 x 32 =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, t = lane & 3;
  // byte offsets inside a panel, for (k-step parity sp, m8-block parity ip):
  // sample row r = 8*(s>>1) + sp + 2t ; column chunk = (ip*4 + g/2) ^ (r & 7)
  uint32_t off[2][2];
#pragma unroll
  for (int sp = 0; sp < 2; ++sp)
#pragma unroll
    for (int ip = 0; ip < 2; ++ip) {
      const int r7 = sp + 2 * t;
      off[sp][ip] = (uint32_t)(r7 * 128 + ((((ip * 4) + (g >> 1)) ^ r7) << 4) + (g & 1) * 8);
    }
  const uint32_t a_cg0 = (uint32_t)(wm * 4) * (kKC * 128);  // column group base of this warp's A rows
  const uint32_t b_cg0 = (uint32_t)(wn * 2) * (kKC * 128);

  int stage = 0;
  uint32_t phase = 0;
  for (long long u = 0;; ++u) {
    long long item; int k_begin, k_end;
    if (!get_unit(u, item, k_begin, k_end)) break;
    int mat, ti, tj;
    decode_item(item, args, mat, ti, tj);
    const bool diag = !args.full && (ti == tj);
    const bool weighted = args.w[mat] != nullptr;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    if (args.ksplit) {
      const int kper = (k_iters + args.n_mats - 1) / args.n_mats;
      k_begin = mat * kper; k_end = min(k_iters, k_begin + kper);
    }
    for (int k = k_begin; k < k_end; ++k) {
      mbar_wait(&full[stage], phase);
      const uint8_t* st = smem + stage * kStageBytes;
      const uint8_t* pa = st + a_cg0;
      const uint8_t* pb = st + (diag ? 0 : kPanelBytes) + b_cg0;
      const double* ws = (const double*)(st + 2 * kPanelBytes);
#pragma unroll
      for (int s = 0; s < kKC / 4; ++s) {
        const int sp = s & 1;
        const uint32_t rbase = (uint32_t)(8 * (s >> 1)) * 128;
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          a[i] = *(const double*)(pa + (uint32_t)(i >> 1) * (kKC * 128) + rbase + off[sp][i & 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          b[j] = *(const double*)(pb + (uint32_t)(j >> 1) * (kKC * 128) + rbase + off[sp][j & 1]);
        if (weighted) {
          const double wv = ws[8 * (s >> 1) + sp + 2 * t];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] *= wv;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }

    // epilogue: S[ti*128 + m][tj*128 + n] += acc   (single writer per tile and launch; the K-slices of a tail item
    // take turns in slice order, so the sum is still formed in a fixed order)
    double* S = args.S[mat];
    const long long row0 = (long long)ti * kBM + wm * 64 + g;
    const long long col0 = (long long)tj * kBN + wn * 32 + 2 * t;
    bool atomic_fallback = false;
    const int tslice = (args.tail_ks && item >= main_items) ? (int)blockIdx.x % args.tail_ks : -1;
    const int tidx = tslice >= 0 ? (int)blockIdx.x / args.tail_ks : 0;
    if (tslice >= 0) {
      volatile int& s_fallback = *(volatile int*)(empty + kStages);   // 8 spare bytes behind the mbarriers
      if (threadIdx.x == 0) {
        int fb = 0;
        if (tslice > 0) {
          const long long t0 = clock64();
          unsigned turn;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(turn) : "l"(&g_tail_turn[tidx]) : "memory");
          } while (turn < (unsigned)tslice && clock64() - t0 < 200000000ll);   // ~0.1 s: never reached when the CTAs are co-resident
          fb = turn < (unsigned)tslice;
        }
        s_fallback = fb;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 consumer warps
      atomic_fallback = s_fallback != 0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double2* p = (double2*)(S + (row0 + i * 8) * args.Pp + col0 + j * 8);
        if (atomic_fallback) {   // a predecessor slice has not shown up: stay correct, give up the fixed order
          atomicAdd(&p->x, args.alpha * acc[i][j][0]);
          atomicAdd(&p->y, args.alpha * acc[i][j][1]);
          continue;
        }
        double2 v = make_double2(0.0, 0.0);
        const double beta = tslice > 0 ? 1.0 : args.beta;   // later slices add onto what the earlier ones wrote
        if (beta != 0.0) { v = *p; v.x *= beta; v.y *= beta; }
        v.x = fma(args.alpha, acc[i][j][0], v.x);
        v.y = fma(args.alpha, acc[i][j][1], v.y);
        *p = v;
      }
    if (tslice >= 0) {
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        if (atomic_fallback) atomicAdd(&g_tail_turn[tidx], 1u);
        else asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&g_tail_turn[tidx]), "r"((unsigned)(tslice + 1)) : "memory");
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D view of the row-major matrix X[n][ldo]: {16 doubles, n rows, Pp/16 column groups}
int make_panel_tensor_map(CUtensorMap* map, const double* X, long long n, long long ldo, int Pp) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(VMCPDE_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {16, (cuuint64_t)n, (cuuint64_t)(Pp / 16)};
  cuuint64_t strides[2] = {(cuuint64_t)ldo * 8, 128};
  cuuint32_t box[3] = {16, (cuuint32_t)kKC, 8};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)X, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VMCPDE_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return 0;
}

static size_t gram_smem_bytes() { return (size_t)kStages * kStageBytes + 2 * kStages * sizeof(uint64_t) + 16; }

// DMMA issue-rate microbenchmark: register-resident chains, no memory traffic.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace vmc

extern "C" __attribute__((visibility("default"))) int vmcpde_gram(const double* O, int64_t n, int64_t ldo, int32_t Pp, int32_t n_mats,
                           const double* const* weights, double* const* S, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(O && S, "vmcpde_gram: null pointer");
  VMC_REQUIRE(Pp > 0 && Pp % 128 == 0, "vmcpde_gram: Pp must be a positive multiple of 128");
  VMC_REQUIRE(ldo >= Pp && ldo % 2 == 0, "vmcpde_gram: ldo must be even and >= Pp");
  VMC_REQUIRE(n >= 0 && n % kKC == 0, "vmcpde_gram: n must be a multiple of 16");
  VMC_REQUIRE(n_mats >= 1 && n_mats <= 4, "vmcpde_gram: n_mats must be in [1,4]");
  VMC_REQUIRE(((uintptr_t)O & 15) == 0, "vmcpde_gram: O must be 16-byte aligned");
  if (n == 0) return 0;
  GramArgs a{};
  for (int m = 0; m < n_mats; ++m) {
    VMC_REQUIRE(S[m] != nullptr, "vmcpde_gram: null output matrix");
    a.S[m] = S[m];
    a.w[m] = weights ? weights[m] : nullptr;
    VMC_REQUIRE(((uintptr_t)a.w[m] & 15) == 0, "vmcpde_gram: weights must be 16-byte aligned");
  }
  a.n_mats = n_mats; a.tiles = Pp / 128; a.tiles_n = Pp / 128; a.Pp = Pp; a.n = n; a.full = 0; a.alpha = 1.0; a.beta = 1.0;
  CUtensorMap map;
  if (int rc = make_panel_tensor_map(&map, O, n, ldo, Pp)) return rc;
  const size_t smem = gram_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    VMC_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long n_items = (long long)n_mats * a.tiles * (a.tiles + 1) / 2;
  int grid = num_sms();
  if (n_items < grid) grid = (int)n_items;
  a.super = 1;
  while ((a.super + 1) * (a.super + 1) * n_mats <= grid) ++a.super;
  if (const char* e_ = getenv("VMCPDE_GRAM_SUPERTILE")) a.super = atoi(e_) > 0 ? atoi(e_) : a.super;
  // Per-item rendezvous of the producers: measured at C3 (3 matrices, n = 2^18) it cuts the DRAM reads from 1.52 TB to
  // 0.41 TB (L2 hit rate 52 % -> 80 %) but costs 1.2 % of time -- the kernel is DMMA bound and the slack that lets fast CTAs
  // run ahead is exactly what a rendezvous removes -- so it is opt-in (VMCPDE_GRAM_WAVE_SYNC=1).
  a.wave_sync = (n_items > grid && n >= 4096 && getenv("VMCPDE_GRAM_WAVE_SYNC")) ? 1 : 0;
  // The last round is usually partly filled (C3: 6240 items on 148 CTAs = 42 rounds + 24 items): split its items along K so
  // that all CTAs share it.  Needs rounds of equal cost (beta = 1 accumulation, long items) to be worth it.
  a.tail_ks = 0;
  {
    const long long rem = n_items % grid;
    if (n_items > grid && rem > 0 && rem * 2 <= grid && n / kKC >= 256 && !getenv("VMCPDE_GRAM_NOTAIL")) {
      int ks = (int)(grid / rem);
      if (ks > 8) ks = 8;
      if (rem <= 256) a.tail_ks = ks;
    }
  }
  if (a.tail_ks) {
    void* tt = nullptr;
    VMC_CUDA_CHECK(cudaGetSymbolAddress(&tt, g_tail_turn));
    VMC_CUDA_CHECK(cudaMemsetAsync(tt, 0, 256 * sizeof(unsigned), (cudaStream_t)stream));
  }
  if (a.wave_sync) {
    void* ctr = nullptr;
    VMC_CUDA_CHECK(cudaGetSymbolAddress(&ctr, g_wave_counter));
    VMC_CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned), (cudaStream_t)stream));
  }
  gram_kernel<<<grid, kGramThreads, smem, (cudaStream_t)stream>>>(map, map, a);
  VMC_LAUNCH_CHECK("gram_kernel");
  return 0;
}

// Upper-triangular tiles of Out = alpha * X^T X + beta * Out (X [K x M] row-major, M multiple of 128, K of 16): the SYRK
// form of the pipeline with a scale, used by the blocked Cholesky's trailing update.
extern "C" __attribute__((visibility("default"))) int vmcpde_syrk_tn(const double* X, int64_t ldx, double* Out, int64_t ldo, int32_t M, int64_t K,
                                                                      double alpha, double beta, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(X && Out, "vmcpde_syrk_tn: null pointer");
  VMC_REQUIRE(M > 0 && M % 128 == 0 && K >= 0 && K % kKC == 0, "vmcpde_syrk_tn: M multiple of 128 and K of 16 required");
  VMC_REQUIRE(ldx >= M && ldo >= M && ldx % 2 == 0 && ldo % 2 == 0 && ldo <= 0x7fffffff, "vmcpde_syrk_tn: bad leading dimensions");
  if (K == 0) return 0;
  GramArgs a{};
  a.S[0] = Out; a.w[0] = nullptr; a.n_mats = 1; a.tiles = M / 128; a.tiles_n = M / 128; a.Pp = (int)ldo; a.n = K; a.full = 0;
  a.alpha = alpha; a.beta = beta;
  CUtensorMap mx;
  if (int rc = make_panel_tensor_map(&mx, X, K, ldx, M)) return rc;
  const size_t smem = gram_smem_bytes();
  VMC_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_items = (long long)a.tiles * (a.tiles + 1) / 2;
  int grid = num_sms();
  if (n_items < grid) grid = (int)n_items;
  a.super = 1;
  while ((a.super + 1) * (a.super + 1) <= grid) ++a.super;
  gram_kernel<<<grid, kGramThreads, smem, (cudaStream_t)stream>>>(mx, mx, a);
  VMC_LAUNCH_CHECK("gram_kernel(syrk_tn)");
  return 0;
}

// General FP64 tensor-core product on the same pipeline: Out[M x N] = alpha * X^T Y + beta * Out with
// X [K x M] (ldx), Y [K x N] (ldy) row-major, i.e. both operands contiguous along the output index
// (the layout every product of the solve stage is arranged to have).  M, N multiples of 128, K of 16.
extern "C" __attribute__((visibility("default"))) int vmcpde_gemm_tn(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Out,
                                                                      int64_t ldo, int32_t M, int32_t N, int64_t K, double alpha, double beta,
                                                                      vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(X && Y && Out, "vmcpde_gemm_tn: null pointer");
  VMC_REQUIRE(M > 0 && N > 0 && M % 128 == 0 && N % 128 == 0 && K >= 0 && K % kKC == 0, "vmcpde_gemm_tn: M, N multiples of 128 and K of 16 required");
  VMC_REQUIRE(ldx >= M && ldy >= N && ldo >= N && ldx % 2 == 0 && ldy % 2 == 0 && ldo % 2 == 0, "vmcpde_gemm_tn: bad leading dimensions");
  VMC_REQUIRE(ldo <= 0x7fffffff, "vmcpde_gemm_tn: ldo too large");
  if (K == 0) return 0;
  GramArgs a{};
  a.S[0] = Out; a.w[0] = nullptr; a.n_mats = 1; a.tiles = M / 128; a.tiles_n = N / 128; a.Pp = (int)ldo; a.n = K; a.full = 1;
  a.super = 1;
  a.alpha = alpha; a.beta = beta;
  CUtensorMap mx, my;
  if (int rc = make_panel_tensor_map(&mx, X, K, ldx, M)) return rc;
  if (int rc = make_panel_tensor_map(&my, Y, K, ldy, N)) return rc;
  const size_t smem = gram_smem_bytes();
  VMC_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_items = (long long)a.tiles * a.tiles_n;
  int grid = num_sms();
  if (n_items < grid) grid = (int)n_items;
  gram_kernel<<<grid, kGramThreads, smem, (cudaStream_t)stream>>>(mx, my, a);
  VMC_LAUNCH_CHECK("gram_kernel(gemm_tn)");
  return 0;
}

// Split-K form of vmcpde_gemm_tn for products with few output tiles and a long contraction (the Y^T V products of the
// back-transformation): slice s of the K range goes to its own output Part + s * M * ldo (alpha = 1, beta = 0), so
// tiles x splits work items fill the GPU; the caller adds the slices (fixed order).  splits in [1, 16].
extern "C" __attribute__((visibility("default"))) int vmcpde_gemm_tn_splitk(const double* X, int64_t ldx, const double* Y, int64_t ldy, double* Part,
                                                                             int64_t ldo, int32_t M, int32_t N, int64_t K, int32_t splits,
                                                                             vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(X && Y && Part, "vmcpde_gemm_tn_splitk: null pointer");
  VMC_REQUIRE(M > 0 && N > 0 && M % 128 == 0 && N % 128 == 0 && K > 0 && K % kKC == 0, "vmcpde_gemm_tn_splitk: M, N multiples of 128 and K of 16 required");
  VMC_REQUIRE(ldx >= M && ldy >= N && ldo >= N && ldx % 2 == 0 && ldy % 2 == 0 && ldo % 2 == 0 && ldo <= 0x7fffffff, "vmcpde_gemm_tn_splitk: bad leading dimensions");
  VMC_REQUIRE(splits >= 1 && splits <= kMaxMats, "vmcpde_gemm_tn_splitk: splits must be in [1, 16]");
  GramArgs a{};
  for (int sidx = 0; sidx < splits; ++sidx) { a.S[sidx] = Part + (size_t)sidx * M * ldo; a.w[sidx] = nullptr; }
  a.n_mats = splits; a.tiles = M / 128; a.tiles_n = N / 128; a.Pp = (int)ldo; a.n = K; a.full = 1; a.ksplit = 1; a.super = 1;
  a.alpha = 1.0; a.beta = 0.0;
  CUtensorMap mx, my;
  if (int rc = make_panel_tensor_map(&mx, X, K, ldx, M)) return rc;
  if (int rc = make_panel_tensor_map(&my, Y, K, ldy, N)) return rc;
  const size_t smem = gram_smem_bytes();
  VMC_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_items = (long long)splits * a.tiles * a.tiles_n;
  int grid = num_sms();
  if (n_items < grid) grid = (int)n_items;
  gram_kernel<<<grid, kGramThreads, smem, (cudaStream_t)stream>>>(mx, my, a);
  VMC_LAUNCH_CHECK("gram_kernel(gemm_tn_splitk)");
  return 0;
}

// One launch of the register-resident DMMA loop (FP64 tensor peak probe: the roofline denominator of the S build).
// `scratch` is 8 bytes of device memory; the caller times the launch with events on `stream` and divides *flops by it.
// No allocation, no synchronisation.
extern "C" __attribute__((visibility("default"))) int vmcpde_dmma_probe(void* scratch, int32_t iters, double* flops, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(scratch && iters > 0, "vmcpde_dmma_probe: bad arguments");
  const int blocks = num_sms() * 4;
  dmma_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)scratch, iters);
  VMC_LAUNCH_CHECK("dmma_peak_kernel");
  if (flops) *flops = (double)blocks * 8 /*warps*/ * iters * 16.0 * 512.0;
  return 0;
}
