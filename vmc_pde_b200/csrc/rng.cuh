// Threefry-2x32-20 in JAX 0.2.18's counter layout (third-party arithmetic the reference relies on:
// sampler.py:26,33,58-60,73; tdvp.py:154-155).  Integer part is bit-exact (Random123 KATs in tests);
// uniform->normal uses CUDA's erfinv (<= a few ulp from XLA's polynomial; stated in DESIGN.md).
#pragma once
#include <cstdint>
#include "flow_core.cuh"

namespace vmc {

VMC_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

VMC_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  x0 += ks[0]; x1 += ks[1];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = rotl32(x1, (i & 1) ? R1[r] : R0[r]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
}

// 64 random bits of element e out of `total` elements (jax _random_bits, bit_width 64):
// block (e, total + e) -> (hi, lo)
VMC_HD uint64_t random_bits64(uint32_t k0, uint32_t k1, uint64_t e, uint64_t total) {
  uint32_t x0 = (uint32_t)e, x1 = (uint32_t)(total + e);
  threefry2x32(k0, k1, x0, x1);
  return ((uint64_t)x0 << 32) | (uint64_t)x1;
}

VMC_HD double bits_to_unit(uint64_t bits) {  // [0,1): mantissa | 1.0, minus 1
  union { uint64_t u; double d; } c;
  c.u = (bits >> 12) | 0x3FF0000000000000ull;
  return c.d - 1.0;
}

#ifdef __CUDACC__
__device__ __forceinline__ double normal_from_bits(uint64_t bits) {
  // _normal_real: u = uniform(minval=nextafter(-1,0), maxval=1); sqrt(2) * erfinv(u)
  const double lo = -0.99999999999999988897769753748;  // nextafter(-1, 0)
  double u = bits_to_unit(bits) * 2.0 + lo;            // (maxval - minval) rounds to 2.0 in float64
  u = fmax(lo, u);
  return 1.4142135623730951 * erfinv(u);
}
#endif

}  // namespace vmc
