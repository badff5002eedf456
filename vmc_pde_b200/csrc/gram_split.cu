// Split-precision weighted Gram  S += O^T diag(w) O  on the 5th-generation tensor cores (tcgen05 / UMMA, TMEM accumulators).
//
// Replaces mpi_wrapper._cov_helper_without_p (mpi_wrapper.py:21-25) as used for the two Gram matrices of a TDVP right-hand
// side that tolerate a stated error: SExp (tdvp.py:47; read only as v^T SExp v against a tolerance, stepper.py:71-79) and the
// SNR covariance C_EO (tdvp.py:68-70; feeds the logged snr).  S0 -- the matrix that is solved -- stays on the FP64 DMMA path
// (gram.cu).  Selected by TDVP(gramPrecision="split"); the default stays FP64 everywhere.
//
// tcgen05.mma has no FP64 kind, so an FP64 operand x (already multiplied by sqrt(w_row)) is split into three bfloat16
// slices x = x1 + x2 + x3 + O(2^-24 |x|) and a logical product a*b becomes the six products a_i b_j with i + j <= 4,
// accumulated in FP32 in TMEM in TWO accumulators (a1 b1, and the five small products) for at most 128 samples, then added
// in FP64 registers by the epilogue warps.  Error of the result, measured on B200: 2.2e-7 relative to sqrt(S_ii S_jj)
// (5.0e-7 with 256-sample chunks: the tensor core truncates its FP32 accumulation, so the error grows linearly with the
// chunk; a NumPy emulation gives 9e-9 for round-to-nearest accumulation) -- the stated tolerance is 1e-6.
//
//  * pre-pass (split_rows_kernel): O [n][ldo] FP64 -> X [3 slices][Pp columns][n_pad samples] bf16, sample index
//    contiguous, so every operand tile is the canonical K-major 64-byte-swizzle UMMA layout and one TMA 2-D box
//    {32 samples, 128 columns} per slice lands it.
//  * main kernel (gram_split_kernel), persistent, one CTA per SM, warp-specialised: warp 0 = TMA producer (4-stage ring of
//    48 KB: A and B tiles x 3 slices x 32 samples), warp 1 = MMA issuer (one elected lane; 12 tcgen05.mma 128x128x16 per
//    stage), warps 2-9 = epilogue (tcgen05.ld of the two 128x128 FP32 accumulators, FP64 add; TMEM double buffered: 4 x 128
//    columns = all 512).  Work items = upper-triangular 128x128 tile pairs in the supertile order of gram.cu; diagonal tiles
//    load one operand.
//  * all mbarrier spins carry a clock time-out that traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdlib>
#include "common.cuh"

namespace vmc {

constexpr int kSpBK = 32;                 // samples per pipeline stage (64 bytes of bf16: one row of the 64-byte swizzle)
constexpr int kSpChunk = 128;             // samples per TMEM accumulation chunk (FP32 accumulation length)
constexpr int kSpTileBytes = 128 * kSpBK * 2;     // one slice of one operand tile: 128 columns x 32 samples x 2 B = 8 KB
constexpr int kSpStageBytes = 6 * kSpTileBytes;   // A x 3 slices, B x 3 slices = 48 KB
constexpr int kSpStages = 4;              // (2 stages of 64 samples left the tensor pipe 64 % busy: a freed slot was refilled
                                          //  only one stage time, 0.8 us, ahead of its use -- less than the TMA round trip)
constexpr int kSpThreads = 320;           // producer warp, MMA warp, 8 epilogue warps
constexpr size_t kSpSmem = 1024 + (size_t)kSpStages * kSpStageBytes + 256;

struct SplitArgs {
  double* S;          // output, leading dimension ldS, upper-triangular tiles are accumulated into
  int ldS;
  int Pp;             // padded columns
  int tiles;          // Pp / 128
  int super;          // supertile edge
  long long n_pad;    // padded samples (multiple of kSpChunk)
  unsigned* fault;    // set to a stage code when a spin timed out
  int collector;      // 1: keep shared A operands in the tensor core's collector across consecutive MMAs
};

__device__ __forceinline__ uint32_t sp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sp_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sp_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool sp_mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(sp_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded spin: ~2 s at 2 GHz, then record where and trap (an error return instead of a hung GPU)
__device__ __forceinline__ void sp_mbar_wait(uint64_t* bar, uint32_t parity, unsigned* fault, unsigned code) {
  if (sp_mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!sp_mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (fault) atomicExch(fault, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void sp_tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(sp_smem_u32(dst)), "l"(map), "r"(sp_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sp_smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 operands, FP32 accumulate), M = N = 128, K = 16.
// A 128x128x16 MMA reads 4 KB of A and 4 KB of B from shared memory in ~68 clocks -- the whole 128 B/clk of the SM, with the
// TMA writes of the next stage on top (ncu: tensor pipe 64 % busy).  The A-collector keeps an A operand inside the tensor core
// across consecutive MMAs that share it (a1 with b1, b2, b3; a2 with b1, b2): 12 KB instead of 24 KB of A per k-step.
// COLL: 0 = plain, 1 = fill (read A, keep it), 2 = use (reuse, keep), 3 = lastuse (reuse, release).
template <int COLL>
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (COLL == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else if constexpr (COLL == 2)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else if constexpr (COLL == 3)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp gets lane (base lane + t)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 64-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte offset = 8 rows x 64 B = 512 B >> 4
// in [32,46), descriptor version 1 (sm_100) in [46,48), layout type 4 = SWIZZLE_64B in [61,64).
__device__ __forceinline__ uint64_t sp_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(8 * kSpBK * 2 / 16) << 32) | (1ull << 46) | (4ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D = F32 (bits [4,6) = 1), A = B = BF16 ([7,10) =
// [10,13) = 1), both K-major ([15], [16] = 0), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kSpIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// work item -> upper-triangular tile pair (ti <= tj), supertile by supertile (same walk as gram.cu::decode_item)
__device__ __forceinline__ void sp_decode(long long p, int tiles, int G, int& ti, int& tj) {
  const int ST = (tiles + G - 1) / G;
  for (int SI = 0; SI < ST; ++SI) {
    const int r0 = SI * G, rn = min(G, tiles - r0);
    long long cnt = (long long)rn * (rn + 1) / 2;
    if (p < cnt) {
      int r = 0;
      while (p >= (long long)(rn - r)) { p -= (rn - r); ++r; }
      ti = r0 + r; tj = r0 + r + (int)p;
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
  ti = tj = 0;
}

// ------------------------------------------------------------------------------------------------ pre-pass
// X[(slice * Pp + c) * n_pad + s] = slice-th bf16 term of sqrt(w[s]) * O[s][c]; zero for s >= n.
// One block: 64 columns x 128 samples (reads 512-byte row pieces, writes 256-byte runs of samples, 4 blocks per SM; with
// 64 samples per block: 10.8 ms at C3, with 256 samples and 99 KB of shared memory per block: 18 ms -- occupancy).
constexpr int kSplitRows = 128;
__global__ void __launch_bounds__(256) split_rows_kernel(const double* __restrict__ O, long long n, long long ldo, int Pp,
                                                         const double* __restrict__ w, long long n_pad,
                                                         __nv_bfloat16* __restrict__ X) {
  extern __shared__ __nv_bfloat16 sp_t[];                      // [3][64][kSplitRows + 2]
  constexpr int LD = kSplitRows + 2;
  const long long s0 = (long long)blockIdx.x * kSplitRows;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;     // tx: column inside the tile, ty: sample sub-row
  __shared__ double sw[kSplitRows];                           // sqrt of the row weights, once per sample (not per element)
  if (threadIdx.x < kSplitRows) {
    const long long s = s0 + threadIdx.x;
    sw[threadIdx.x] = (s < n) ? (w ? sqrt(w[s]) : 1.0) : 0.0;
  }
  __syncthreads();
#pragma unroll 4
  for (int i = ty; i < kSplitRows; i += 4) {
    const long long s = s0 + i;
    double x = 0.0;
    if (s < n) x = O[s * ldo + c0 + tx] * sw[i];
    const __nv_bfloat16 a1 = __float2bfloat16_rn((float)x);
    const double r1 = x - (double)__bfloat162float(a1);
    const __nv_bfloat16 a2 = __float2bfloat16_rn((float)r1);
    const double r2 = r1 - (double)__bfloat162float(a2);
    const __nv_bfloat16 a3 = __float2bfloat16_rn((float)r2);
    sp_t[(0 * 64 + tx) * LD + i] = a1; sp_t[(1 * 64 + tx) * LD + i] = a2; sp_t[(2 * 64 + tx) * LD + i] = a3;
  }
  __syncthreads();
  // 3 x 64 rows of kSplitRows samples: one warp per row, 2 samples (4 B) per lane and 64-sample piece
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < 192; row += 8) {
    const int c = row & 63, sl = row >> 6;
    __nv_bfloat16* dst = X + ((long long)sl * Pp + c0 + c) * n_pad + s0;
#pragma unroll
    for (int j = 0; j < kSplitRows / 64; ++j) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(&sp_t[row * LD + 64 * j + 2 * lane]);
      *reinterpret_cast<uint32_t*>(dst + 64 * j + 2 * lane) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ main kernel
__global__ void __launch_bounds__(kSpThreads, 1)
gram_split_kernel(const __grid_constant__ CUtensorMap tmap, const SplitArgs a) {
  extern __shared__ uint8_t sp_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)sp_raw + 1023) & ~(uintptr_t)1023);     // swizzle-128B tiles need 1 KB alignment
  uint64_t* bars = (uint64_t*)(smem + (size_t)kSpStages * kSpStageBytes);
  uint64_t* full = bars;                    // [kSpStages] TMA -> MMA
  uint64_t* empty = bars + kSpStages;       // [kSpStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kSpStages;   // [2] MMA -> epilogue (TMEM buffer ready)
  uint64_t* tempty = tfull + 2;             // [2] epilogue -> MMA (TMEM buffer drained)
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSpStages; ++i) { sp_mbar_init(&full[i], 1); sp_mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { sp_mbar_init(&tfull[i], 1); sp_mbar_init(&tempty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) {   // the whole of TMEM: 2 buffers x (hi, lo) x 128 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sp_smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_items = (long long)a.tiles * (a.tiles + 1) / 2;
  const int nkb = (int)(a.n_pad / kSpBK);
  const int nchunks = (int)(a.n_pad / kSpChunk);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        int ti, tj;
        sp_decode(item, a.tiles, a.super, ti, tj);
        const bool diag = ti == tj;
        for (int kb = 0; kb < nkb; ++kb) {
          sp_mbar_wait(&empty[stage], phase ^ 1, a.fault, 1);
          uint8_t* sA = smem + (size_t)stage * kSpStageBytes;
          uint8_t* sB = sA + 3 * kSpTileBytes;
          sp_mbar_expect_tx(&full[stage], (diag ? 3 : 6) * kSpTileBytes);
#pragma unroll
          for (int sl = 0; sl < 3; ++sl) sp_tma_2d(sA + sl * kSpTileBytes, &tmap, &full[stage], kb * kSpBK, sl * a.Pp + ti * 128);
          if (!diag) {
#pragma unroll
            for (int sl = 0; sl < 3; ++sl) sp_tma_2d(sB + sl * kSpTileBytes, &tmap, &full[stage], kb * kSpBK, sl * a.Pp + tj * 128);
          }
          if (++stage == kSpStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      uint32_t cidx = 0;   // chunk counter over the whole kernel: TMEM buffer = cidx & 1
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        int ti, tj;
        sp_decode(item, a.tiles, a.super, ti, tj);
        const bool diag = ti == tj;
        for (int ch = 0; ch < nchunks; ++ch, ++cidx) {
          const uint32_t buf = cidx & 1, tphase = (cidx >> 1) & 1;
          sp_mbar_wait(&tempty[buf], tphase ^ 1, a.fault, 2);
          tc_fence_after();
          const uint32_t d_hi = tmem_base + buf * 256, d_lo = d_hi + 128;
          for (int kb4 = 0; kb4 < kSpChunk / kSpBK; ++kb4) {
            sp_mbar_wait(&full[stage], phase, a.fault, 3);
            tc_fence_after();
            const uint32_t sA = sp_smem_u32(smem + (size_t)stage * kSpStageBytes);
            const uint32_t sB = diag ? sA : sA + 3 * kSpTileBytes;
#pragma unroll
            for (int ks = 0; ks < kSpBK / 16; ++ks) {
              const uint32_t ko = ks * 32;   // 16 bf16 along K inside the swizzle row
              const uint64_t a1 = sp_desc(sA + ko), a2 = sp_desc(sA + kSpTileBytes + ko), a3 = sp_desc(sA + 2 * kSpTileBytes + ko);
              const uint64_t b1 = sp_desc(sB + ko), b2 = sp_desc(sB + kSpTileBytes + ko), b3 = sp_desc(sB + 2 * kSpTileBytes + ko);
              const uint32_t first = (kb4 == 0 && ks == 0) ? 0u : 1u;
              if (a.collector) {
                tc_mma_bf16<1>(d_hi, a1, b1, kSpIdesc, first);
                tc_mma_bf16<2>(d_lo, a1, b2, kSpIdesc, first);
                tc_mma_bf16<3>(d_lo, a1, b3, kSpIdesc, 1u);
                tc_mma_bf16<1>(d_lo, a2, b1, kSpIdesc, 1u);
                tc_mma_bf16<3>(d_lo, a2, b2, kSpIdesc, 1u);
                tc_mma_bf16<0>(d_lo, a3, b1, kSpIdesc, 1u);
              } else {
                tc_mma_bf16<0>(d_hi, a1, b1, kSpIdesc, first);
                tc_mma_bf16<0>(d_lo, a1, b2, kSpIdesc, first);
                tc_mma_bf16<0>(d_lo, a2, b1, kSpIdesc, 1u);
                tc_mma_bf16<0>(d_lo, a1, b3, kSpIdesc, 1u);
                tc_mma_bf16<0>(d_lo, a2, b2, kSpIdesc, 1u);
                tc_mma_bf16<0>(d_lo, a3, b1, kSpIdesc, 1u);
              }
            }
            tc_commit(&empty[stage]);         // frees the smem slot when these MMAs have read it
            if (++stage == kSpStages) { stage = 0; phase ^= 1; }
          }
          tc_commit(&tfull[buf]);             // accumulators of this chunk complete
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> FP64 registers; at the end of an item -> global =====
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;            // column half of the tile
    const int row = q * 32 + lane;
    uint32_t cidx = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      int ti, tj;
      sp_decode(item, a.tiles, a.super, ti, tj);
      double acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.0;
      for (int ch = 0; ch < nchunks; ++ch, ++cidx) {
        const uint32_t buf = cidx & 1, tphase = (cidx >> 1) & 1;
        sp_mbar_wait(&tfull[buf], tphase, a.fault, 4);
        tc_fence_after();
        const uint32_t t_hi = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + h * 64;
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          uint32_t hi[16], lo[16];
          tc_ld16(t_hi + part * 16, hi);
          tc_ld16(t_hi + 128 + part * 16, lo);
          tc_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k) acc[part * 16 + k] += (double)__uint_as_float(hi[k]) + (double)__uint_as_float(lo[k]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) sp_mbar_arrive(&tempty[buf]);
      }
      double* out = a.S + (size_t)(ti * 128 + row) * a.ldS + tj * 128 + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        double2 v = *reinterpret_cast<double2*>(out + i);
        v.x += acc[i]; v.y += acc[i + 1];
        *reinterpret_cast<double2*>(out + i) = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

typedef CUresult (*SpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static SpEncodeTiledFn sp_encode_fn() {
  static SpEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (SpEncodeTiledFn)p;
  }
  return fn;
}

__device__ unsigned g_split_fault;

}  // namespace vmc

using namespace vmc;

static long long sp_pad(long long n) { return (n + kSpChunk - 1) / kSpChunk * kSpChunk; }

extern "C" __attribute__((visibility("default"))) int vmcpde_gram_split_workspace_bytes(int64_t n, int32_t Pp, size_t* bytes) {
  VMC_REQUIRE(bytes && n >= 0 && Pp > 0 && Pp % 128 == 0, "vmcpde_gram_split_workspace_bytes: bad arguments");
  *bytes = (size_t)3 * Pp * sp_pad(n) * 2 + 1024;
  return 0;
}

// S (ldS = Pp, upper-triangular 128x128 tiles) += O^T diag(w) O with bf16 x 3 split operands on tcgen05 (see file header).
// O [n][ldo] FP64 row-major (centred), w[n] >= 0 row weights or NULL.  Stated tolerance: 1e-6 relative to sqrt(S_ii S_jj).
extern "C" __attribute__((visibility("default"))) int vmcpde_gram_split(const double* O, int64_t n, int64_t ldo, int32_t Pp, const double* w,
                                                                       double* S, void* workspace, size_t workspace_bytes,
                                                                       vmcpde_stream stream) {
  VMC_REQUIRE(O && S && workspace, "vmcpde_gram_split: null pointer");
  VMC_REQUIRE(Pp > 0 && Pp % 128 == 0 && ldo >= Pp && n >= 0, "vmcpde_gram_split: bad dimensions");
  if (n == 0) return 0;
  const long long n_pad = sp_pad(n);
  size_t need = 0;
  vmcpde_gram_split_workspace_bytes(n, Pp, &need);
  VMC_REQUIRE(workspace_bytes >= need, "vmcpde_gram_split: workspace too small");
  VMC_REQUIRE((long long)3 * Pp < (1ll << 31) && n_pad < (1ll << 31), "vmcpde_gram_split: problem too large for one tensor map");
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16* X = (__nv_bfloat16*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const size_t split_smem = (size_t)3 * 64 * (kSplitRows + 2) * sizeof(__nv_bfloat16);
  static bool split_attr = false;
  if (!split_attr) {
    VMC_CUDA_CHECK(cudaFuncSetAttribute(split_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split_smem));
    split_attr = true;
  }
  split_rows_kernel<<<dim3((unsigned)(n_pad / kSplitRows), Pp / 64), 256, split_smem, s>>>(O, n, ldo, Pp, w, n_pad, X);
  VMC_LAUNCH_CHECK("split_rows_kernel");
  SpEncodeTiledFn enc = sp_encode_fn();
  if (!enc) return set_error(VMCPDE_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)n_pad, (cuuint64_t)3 * Pp};
  cuuint64_t strides[1] = {(cuuint64_t)n_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)kSpBK, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VMCPDE_ECUDA, "cuTensorMapEncodeTiled (bf16 slices) failed with code " + std::to_string((int)r));
  static bool attr_set = false;
  if (!attr_set) {
    VMC_CUDA_CHECK(cudaFuncSetAttribute(gram_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpSmem));
    attr_set = true;
  }
  SplitArgs a{};
  a.S = S; a.ldS = Pp; a.Pp = Pp; a.tiles = Pp / 128; a.n_pad = n_pad;
  a.collector = getenv("VMCPDE_SPLIT_NO_COLLECTOR") ? 0 : 1;
  VMC_CUDA_CHECK(cudaGetSymbolAddress((void**)&a.fault, g_split_fault));
  const long long n_items = (long long)a.tiles * (a.tiles + 1) / 2;
  int grid = num_sms();
  if (n_items < grid) grid = (int)n_items;
  a.super = 1;
  while ((a.super + 1) * (a.super + 1) <= grid) ++a.super;
  gram_split_kernel<<<grid, kSpThreads, kSpSmem, s>>>(map, a);
  VMC_LAUNCH_CHECK("gram_split_kernel");
  return 0;
}
