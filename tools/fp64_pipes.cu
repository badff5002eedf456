// Microbenchmark: issue rates of the FP64 pipes on sm_100a (DMMA.8x8x4 vs DFMA, alone and mixed).
// Evidence for the Gram kernel's roofline denominator; results are committed under profiles/.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC, int MODE>  // MODE 0: DMMA only, 1: DFMA only, 2: per-warp mix (NACC dmma + 8*NACC dfma interleaved), 3: warp-specialised mix
__global__ void k(double* out, long long* cyc, int iters) {
  double acc[NACC][2], f[NACC][8];
  for (int i = 0; i < NACC; ++i) { acc[i][0] = acc[i][1] = 0; for (int j = 0; j < 8; ++j) f[i][j] = 0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2 || (MODE == 3 && (warp & 1) == 0)) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(acc[i][0], acc[i][1], a, b);
    }
    if (MODE == 1 || MODE == 2 || (MODE == 3 && (warp & 1) == 1)) {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i][j] = fma(a, b, f[i][j]);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NACC; ++i) { s += acc[i][0] + acc[i][1]; for (int j = 0; j < 8; ++j) s += f[i][j]; }
  if (s == 1.2345) out[0] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
}
template <int NACC, int MODE>
void run(const char* name, int warps, int blocks_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8 * 4096 * 64);
  const int iters = 4000, blocks = sms * blocks_per_sm;
  k<NACC, MODE><<<blocks, warps * 32>>>(d, c, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<NACC, MODE><<<blocks, warps * 32>>>(d, c, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static long long h[4096 * 64]; cudaMemcpy(h, c, 8 * blocks * warps, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < blocks * warps; ++i) if (h[i] > mx) mx = h[i];
  // FMA lane-ops issued per SM
  double dmma_warps = (MODE == 0 || MODE == 2) ? warps : (MODE == 3 ? warps / 2 : 0);
  double dfma_warps = (MODE == 1 || MODE == 2) ? warps : (MODE == 3 ? warps / 2 : 0);
  double fma_per_sm = blocks_per_sm * (double)iters * NACC * (dmma_warps * 256.0 + dfma_warps * 8 * 32.0);
  double tot = fma_per_sm * sms * 2.0;
  printf("%-28s warps/CTA=%2d CTAs/SM=%d NACC=%2d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM (max warp cycles %lld)  eff clk %.0f MHz\n", name, warps, blocks_per_sm,
         NACC, ms, tot / ms * 1e-9, fma_per_sm / (double)mx, mx, mx / (ms * 1e3));
  cudaFree(d); cudaFree(c);
}
int main() {
  run<8, 0>("DMMA only", 4, 1); run<8, 0>("DMMA only", 8, 1); run<16, 0>("DMMA only", 8, 1); run<16, 0>("DMMA only", 16, 1); run<16, 0>("DMMA only", 8, 4);
  run<4, 1>("DFMA only", 4, 1); run<4, 1>("DFMA only", 8, 1); run<8, 1>("DFMA only", 16, 1); run<8, 1>("DFMA only", 8, 4);
  run<8, 2>("per-warp mix 1:8", 8, 1); run<8, 2>("per-warp mix 1:8", 16, 1);
  run<8, 3>("warp-specialised mix", 8, 1); run<8, 3>("warp-specialised mix", 16, 1);
  return 0;
}
