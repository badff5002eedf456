// tcgen05.mma issue-rate probe: one CTA per SM issues MMAs from (zeroed) shared memory, no loads.
// usage: utc_rate <kind: 0 bf16 128x128x16 | 1 i8 128x64x32 | 2 i8 128x128x32 | 3 i8 128x256x32 | 4 bf16 128x256x16> <collector 0/1> [iters]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo16, uint64_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)sbo16 << 32) | (1ull << 46) | (layout << 61);
}
template <int KIND, int COLL>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (KIND == 0 || KIND == 4) {
    if constexpr (COLL == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (COLL == 2) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  } else {
    if constexpr (COLL == 1) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (COLL == 2) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  }
}
template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int coll, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  constexpr int N = KIND == 0 ? 128 : (KIND == 1 ? 64 : (KIND == 2 ? 128 : 256));
  constexpr bool BF = KIND == 0 || KIND == 4;
  // K-major rows of 32 bytes (one UMMA K step: 16 bf16 or 32 int8), 32-byte swizzle: 8-row atoms of 256 B
  const uint32_t idesc = BF ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                                   : ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      // 8 A tiles (4 KB each) and 8 B tiles: slices p, q; 36 products p + q <= 7 into 8 accumulators (or 6 products for bf16)
      if (BF) {
        for (int p = 0; p < 3; ++p)
          for (int q = 0; q < 3 - p; ++q) {
            const uint64_t a = desc(su32(smem) + p * 4096, 16, 6), b = desc(su32(smem) + 32768 + q * (N * 32), 16, 6);
            const uint32_t d = tm + (p + q == 0 ? 0 : N);
            if (coll && q == 0 && p < 2) mma<KIND, 1>(d, a, b, idesc, 1); else if (coll && q + 1 < 3 - p) mma<KIND, 2>(d, a, b, idesc, 1); else mma<KIND, 0>(d, a, b, idesc, 1);
          }
      } else {
        constexpr int NACC = 512 / N;    // accumulators that fit
        for (int p = 0; p < 8; ++p)
          for (int q = 0; q < 8 - p; ++q) {
            const uint64_t a = desc(su32(smem) + p * 4096, 16, 6), b = desc(su32(smem) + 32768 + q * (N * 32), 16, 6);
            const uint32_t d = tm + ((p + q) % NACC) * N;
            if (coll && q == 0 && p < 7) mma<KIND, 1>(d, a, b, idesc, 1); else if (coll && q + 1 < 8 - p) mma<KIND, 2>(d, a, b, idesc, 1); else mma<KIND, 0>(d, a, b, idesc, 1);
          }
      }
      if ((it & 15) == 15 || it + 1 == iters) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(su32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
      }
    }
    if (blockIdx.x == 0) *cycles = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
int main(int argc, char** argv) {
  const int kind = argc > 1 ? atoi(argv[1]) : 1, coll = argc > 2 ? atoi(argv[2]) : 1, iters = argc > 3 ? atoi(argv[3]) : 20000;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* cyc; cudaMalloc(&cyc, 8);
  const size_t smem = 100 * 1024;
  auto launch = [&](int it) {
    if (kind == 0) { cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate_kernel<0><<<sms, 128, smem>>>(it, coll, cyc); }
    if (kind == 1) { cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate_kernel<1><<<sms, 128, smem>>>(it, coll, cyc); }
    if (kind == 2) { cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate_kernel<2><<<sms, 128, smem>>>(it, coll, cyc); }
    if (kind == 3) { cudaFuncSetAttribute(rate_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate_kernel<3><<<sms, 128, smem>>>(it, coll, cyc); }
    if (kind == 4) { cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate_kernel<4><<<sms, 128, smem>>>(it, coll, cyc); }
  };
  launch(2000);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const int N = kind == 0 ? 128 : (kind == 1 ? 64 : (kind == 2 ? 128 : 256));
  const bool bf = kind == 0 || kind == 4;
  const double mmas = (bf ? 6.0 : 36.0) * iters, ops = mmas * 2.0 * 128 * N * (bf ? 16 : 32) * sms;
  printf("{\"kind\": %d, \"N\": %d, \"collector\": %d, \"ms\": %.3f, \"clk_per_mma\": %.1f, \"tops\": %.1f, \"err\": \"%s\"}\n", kind, N, coll, ms, (double)h / mmas,
         ops / (ms * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
