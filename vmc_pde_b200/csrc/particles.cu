// Stochastic particle integrator, the reference's independent check of the density dynamics
// (exact_dyn.py:22-84): every particle advances by the 4-stage scheme of integrate_single_coord with noise drawn from
// its own key (split(key, N)[i] -> split(.., 4)), for the phase-space Fokker-Planck equation (Hamiltonian advection,
// momentum diffusion and damping) or plain diffusion.  One thread per particle; RNG in JAX's counter layout (rng.cuh).
#include <cstdint>
#include "common.cuh"
#include "rng.cuh"

namespace vmc {

constexpr int kPartMaxDim = 16;

struct PartParams {
  int d, update, field;
  double dt, D, m, omega, lam, T, gamma, t;
};

template <int D>
__device__ __forceinline__ void velocity(const PartParams& p, const double (&x)[D], double (&v)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) v[i] = 0.0;
  if (p.field == 0) {         // exact_dyn.py:31-47 (uncoupled): J grad H on interleaved (x, p)
#pragma unroll
    for (int i = 0; i + 1 < D; i += 2) {
      v[i] = x[i + 1] / p.m;
      v[i + 1] = -(p.m * p.omega * p.omega * x[i] + 4.0 * p.lam * x[i] * x[i] * x[i]);
    }
  } else if (p.field == 1) {  // exact_dyn.py:50-53
    if (D >= 2) {
      const double pi = 3.14159265358979323846, c = cos(pi * p.t / p.T);
      const double sx = sin(pi * x[0]), sy = sin(pi * x[D >= 2 ? 1 : 0]);
      v[0] = -sx * sx * sin(2.0 * pi * x[D >= 2 ? 1 : 0]) * c;
      v[D >= 2 ? 1 : 0] = sy * sy * sin(2.0 * pi * x[0]) * c;
    }
  }
}

// update_fun_phaseSpace / update_fun_Diff (exact_dyn.py:56-67) with the stage's own dt and key
template <int D>
__device__ __forceinline__ void stage(const PartParams& p, const double (&x)[D], double dts, uint32_t s0, uint32_t s1, double (&k)[D]) {
  double xi[D];
#pragma unroll
  for (int a = 0; a < D; ++a) xi[a] = normal_from_bits(random_bits64(s0, s1, (uint64_t)a, (uint64_t)D));
  if (p.update == 1) {
    const double f = p.D * sqrt(2.0 / dts);
#pragma unroll
    for (int a = 0; a < D; ++a) k[a] = f * xi[a];
    return;
  }
  double v[D];
  velocity<D>(p, x, v);
  const double f = sqrt(2.0 * p.m * p.gamma * p.T / dts);
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const double mask = (a & 1) ? 1.0 : 0.0;
    k[a] = v[a] + f * xi[a] * mask + (-p.gamma * x[a]) * mask;
  }
}

// everything is indexed at compile time (registers): one instantiation per dimension
template <int D>
__global__ void __launch_bounds__(128) particles_kernel(double* __restrict__ coords, long long n, PartParams p, uint32_t key0, uint32_t key1) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // keys = split(key, n): word e of the 2 n word stream is the first word of block (e, n + e) for e < n, else the second
  // word of block (e - n, e); the particle key is words (2 i, 2 i + 1)
  uint32_t pk[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint64_t e = 2ull * (uint64_t)i + q, h = (uint64_t)n;
    uint32_t x0 = (uint32_t)(e < h ? e : e - h), x1 = (uint32_t)(e < h ? h + e : e);
    threefry2x32(key0, key1, x0, x1);
    pk[q] = e < h ? x0 : x1;
  }
  // split(particle key, 4): the four blocks (c, 4 + c) give words c (first) and 4 + c (second) of the 8-word stream
  uint32_t sk[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t x0 = (uint32_t)c, x1 = (uint32_t)(4 + c);
    threefry2x32(pk[0], pk[1], x0, x1);
    sk[c] = x0;
    sk[4 + c] = x1;
  }
  double x[D], y[D], k[D], acc[D];
#pragma unroll
  for (int a = 0; a < D; ++a) x[a] = coords[i * D + a];
  const double dt = p.dt;
  // x + dt (k1 + 2 k2 + 2 k3 + k4) / 6, summed left to right as the reference writes it (exact_dyn.py:76)
  stage<D>(p, x, dt / 6, sk[0], sk[1], k);
#pragma unroll
  for (int a = 0; a < D; ++a) { acc[a] = k[a]; y[a] = x[a] + dt * 0.5 * k[a]; }
  stage<D>(p, y, dt / 3, sk[2], sk[3], k);
#pragma unroll
  for (int a = 0; a < D; ++a) { acc[a] = acc[a] + 2.0 * k[a]; y[a] = x[a] + dt * 0.5 * k[a]; }
  stage<D>(p, y, dt / 3, sk[4], sk[5], k);
#pragma unroll
  for (int a = 0; a < D; ++a) { acc[a] = acc[a] + 2.0 * k[a]; y[a] = x[a] + dt * k[a]; }
  stage<D>(p, y, dt / 6, sk[6], sk[7], k);
#pragma unroll
  for (int a = 0; a < D; ++a) coords[i * D + a] = x[a] + dt * (acc[a] + k[a]) / 6.0;
}

template <int D>
static void launch_particles(double* coords, long long n, const PartParams& p, uint32_t key0, uint32_t key1, cudaStream_t s) {
  particles_kernel<D><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(coords, n, p, key0, key1);
}

}  // namespace vmc

// One step of exact_dyn.integrate (exact_dyn.py:70-84) for all n particles, in place.
// update: 0 = update_fun_phaseSpace, 1 = update_fun_Diff; field: 0 = _velocity_field_hamiltonian (uncoupled),
// 1 = _velocity_field_fluiddynpaper, -1 = none.  eq supplies D, m, omega, lam, T, gamma, t (its mode is ignored).
extern "C" __attribute__((visibility("default"))) int vmcpde_particles_step(double* coords, int64_t n, int32_t d, double dt, int32_t update,
                                                                           int32_t field, const vmcpde_equation* eq, uint32_t key0,
                                                                           uint32_t key1, vmcpde_stream stream) {
  using namespace vmc;
  VMC_REQUIRE(coords && eq, "vmcpde_particles_step: null pointer");
  VMC_REQUIRE(d >= 1 && d <= kPartMaxDim, "vmcpde_particles_step: dimension must be in [1, 16]");
  VMC_REQUIRE(update == 0 || update == 1, "vmcpde_particles_step: update must be 0 (phase space) or 1 (diffusion)");
  VMC_REQUIRE(field >= -1 && field <= 1 && (field != 1 || d >= 2), "vmcpde_particles_step: bad velocity field");
  VMC_REQUIRE(dt > 0.0, "vmcpde_particles_step: dt must be positive");
  if (n <= 0) return 0;
  PartParams p{d, update, field, dt, eq->D, eq->m, eq->omega, eq->lam, eq->T, eq->gamma, eq->t};
  cudaStream_t s = (cudaStream_t)stream;
  switch (d) {
#define VMC_PART_CASE(Dv) case Dv: launch_particles<Dv>(coords, n, p, key0, key1, s); break;
    VMC_PART_CASE(1) VMC_PART_CASE(2) VMC_PART_CASE(3) VMC_PART_CASE(4) VMC_PART_CASE(5) VMC_PART_CASE(6) VMC_PART_CASE(7)
    VMC_PART_CASE(8) VMC_PART_CASE(9) VMC_PART_CASE(10) VMC_PART_CASE(11) VMC_PART_CASE(12) VMC_PART_CASE(13)
    VMC_PART_CASE(14) VMC_PART_CASE(15) VMC_PART_CASE(16)
#undef VMC_PART_CASE
  }
  VMC_LAUNCH_CHECK("particles_kernel");
  return 0;
}
