// NVLink flag ping-pong between two GPUs (one process per GPU, peer pointers from torch symmetric memory):
// measures the round-trip latency a distributed panel factorisation would pay per cross-GPU exchange.
// Test tooling only (not part of libvmcpde.so).  Every spin has a clock-based timeout.
#include <cuda_runtime.h>
#include <cstdint>
extern "C" {
__global__ void pingpong_kernel(volatile unsigned long long* local_flag, unsigned long long* peer_flag, int rank, int iters,
                                long long timeout_cycles, unsigned long long* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long t0 = clock64();
  unsigned long long ok = 1;
  for (int i = 1; i <= iters && ok; ++i) {
    if (rank == 0) {
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"((unsigned long long)i) : "memory");
      const long long ts = clock64();
      while (*local_flag < (unsigned long long)i) { if (clock64() - ts > timeout_cycles) { ok = 0; break; } }
    } else {
      const long long ts = clock64();
      while (*local_flag < (unsigned long long)i) { if (clock64() - ts > timeout_cycles) { ok = 0; break; } }
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"((unsigned long long)i) : "memory");
    }
  }
  out[0] = (unsigned long long)(clock64() - t0);
  out[1] = ok;
}
int pingpong(void* local_flag, void* peer_flag, int rank, int iters, long long timeout_cycles, void* out, void* stream) {
  pingpong_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((volatile unsigned long long*)local_flag, (unsigned long long*)peer_flag, rank,
                                                      iters, timeout_cycles, (unsigned long long*)out);
  return (int)cudaGetLastError();
}
}
