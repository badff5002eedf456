// Torch-free caller of the C-ABI (include/vmcpde.h): cudaMalloc'ed buffers, the default stream, one right-hand side of the
// 'mwe' configuration (main.py:38: d = 2, depth 4, intmediate (1,), P = 37, diffusion) through
//   vmcpde_sample -> vmcpde_local_terms -> vmcpde_moments1 -> vmcpde_center_force -> vmcpde_gram -> vmcpde_sym_finalize
//   -> vmcpde_eigh -> vmcpde_solve_tail,
// plus the blocked eigensolver in its two-call form on a 512 x 512 Gram and the tcgen05 split Gram.  Demonstrates "raw
// pointers + stream, caller-owned workspaces" without any framework.  Built by __graft_entry__.build(), run by
// tests/test_cabi.py::test_torch_free_c_caller (gpu).  Prints "CABI_SMOKE OK ..." and exits 0 on success.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/vmcpde.h"

#define CK(x) do { int rc_ = (x); if (rc_) { std::printf("FAIL %s -> %d: %s\n", #x, rc_, vmcpde_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA FAIL %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <class T> static T* dalloc(size_t n) { void* p = nullptr; if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) std::exit(2); cudaMemset(p, 0, n * sizeof(T)); return (T*)p; }

int main() {
  if (vmcpde_version() != VMCPDE_VERSION) { std::printf("version mismatch\n"); return 1; }
  const int d = 2, depth = 4, N = 4096;
  int32_t up[depth] = {0, 1, 0, 1}, down[depth] = {1, 0, 1, 0};
  double offset[d] = {0.0, 0.0};
  vmcpde_flow_config cfg{d, depth, 1, 1, VMCPDE_NO_ADD, VMCPDE_GAUSS, up, down, offset, nullptr};
  vmcpde_flow* flow = nullptr;
  CK(vmcpde_flow_create(&cfg, &flow));
  const int P = vmcpde_flow_num_params(flow), Pp = vmcpde_padded_params(P);
  if (P != 37 || Pp != 128) { std::printf("unexpected parameter count %d / %d\n", P, Pp); return 1; }
  // parameters: latent part zero, kernels small and deterministic
  std::vector<double> theta(P, 0.0);
  std::vector<int32_t> offs(4 + depth);
  CK(vmcpde_flow_param_offsets(flow, offs.data()));
  for (int i = offs[4]; i < P; ++i) theta[i] = 0.3 * std::sin(1.7 * i + 0.3);
  double* d_theta = dalloc<double>(P);
  CU(cudaMemcpy(d_theta, theta.data(), P * 8, cudaMemcpyHostToDevice));
  cudaStream_t s = 0;
  double *x = dalloc<double>((size_t)N * d), *lp = dalloc<double>(N), *E = dalloc<double>(N), *O = dalloc<double>((size_t)N * Pp);
  CK(vmcpde_sample(flow, d_theta, 0u, 42u, 0, N, N, nullptr, x, lp, nullptr, s));
  vmcpde_equation eq{}; eq.mode = VMCPDE_DIFFUSION; eq.D = 1.0;
  CK(vmcpde_local_terms(flow, d_theta, x, N, &eq, E, lp, nullptr, nullptr, O, Pp, s));
  size_t mws_b = 0; CK(vmcpde_moments_workspace_bytes(N, Pp, &mws_b));
  void* mws = dalloc<char>(mws_b);
  double* sums = dalloc<double>(4 + Pp);
  CK(vmcpde_moments1(E, lp, O, N, Pp, sums, mws, mws_b, s));
  std::vector<double> h_sums(4 + Pp);
  CU(cudaMemcpyAsync(h_sums.data(), sums, (4 + Pp) * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  std::vector<double> meanO(Pp);
  for (int p = 0; p < Pp; ++p) meanO[p] = h_sums[4 + p] / N;
  double* d_meanO = dalloc<double>(Pp);
  CU(cudaMemcpy(d_meanO, meanO.data(), Pp * 8, cudaMemcpyHostToDevice));
  double *dE = dalloc<double>(N), *wE = dalloc<double>(N), *wLp = dalloc<double>(N), *Fsum = dalloc<double>(Pp), *var = dalloc<double>(8);
  CK(vmcpde_center_force(O, N, Pp, d_meanO, E, lp, h_sums[0] / N, dE, wE, wLp, Fsum, var, mws, mws_b, s));
  double *S0 = dalloc<double>((size_t)Pp * Pp), *SExp = dalloc<double>((size_t)Pp * Pp), *CEO = dalloc<double>((size_t)Pp * Pp);
  const double* weights[3] = {nullptr, wLp, wE};
  double* mats[3] = {S0, SExp, CEO};
  CK(vmcpde_gram(O, N, Pp, Pp, 3, weights, mats, s));
  for (double* m : mats) CK(vmcpde_sym_finalize(m, Pp, 1.0 / N, s));
  // the same SExp on the tcgen05 split path
  double* SExp2 = dalloc<double>((size_t)Pp * Pp);
  size_t sp_b = 0; CK(vmcpde_gram_split_workspace_bytes(N, Pp, &sp_b));
  void* spws = dalloc<char>(sp_b);
  CK(vmcpde_gram_split(O, N, Pp, Pp, wLp, SExp2, spws, sp_b, s));
  CK(vmcpde_sym_finalize(SExp2, Pp, 1.0 / N, s));
  // F = Fsum / N on the host (P doubles), solve
  std::vector<double> F(Pp);
  CU(cudaMemcpyAsync(F.data(), Fsum, Pp * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (auto& v : F) v /= N;
  double* d_F = dalloc<double>(Pp);
  CU(cudaMemcpy(d_F, F.data(), Pp * 8, cudaMemcpyHostToDevice));
  double *work = dalloc<double>((size_t)Pp * Pp), *ev = dalloc<double>(Pp), *VT = dalloc<double>((size_t)Pp * Pp);
  CU(cudaMemcpyAsync(work, S0, (size_t)Pp * Pp * 8, cudaMemcpyDeviceToDevice, s));
  size_t ews_b = 0, tws_b = 0;
  CK(vmcpde_eigh_workspace_bytes(P, Pp, &ews_b));
  CK(vmcpde_solve_tail_workspace_bytes(P, Pp, &tws_b));
  const size_t ws_b = ews_b > tws_b ? ews_b : tws_b;
  void* ws = dalloc<char>(ws_b);
  CK(vmcpde_eigh(work, P, Pp, ev, VT, ws, ws_b, s));
  double *VtF = dalloc<double>(Pp), *rho = dalloc<double>(Pp), *snr = dalloc<double>(Pp), *inv = dalloc<double>(Pp), *upd = dalloc<double>(Pp), *scal = dalloc<double>(2);
  CK(vmcpde_solve_tail(ev, VT, P, Pp, d_F, S0, S0, CEO, (double)N, 1e-11, 2.0, 0, h_sums[2] / N, VtF, rho, snr, inv, upd, scal, ws, ws_b, s));
  std::vector<double> h_scal(2), h_ev(Pp), h_a((size_t)Pp * Pp), h_b((size_t)Pp * Pp);
  CU(cudaMemcpyAsync(h_scal.data(), scal, 16, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_ev.data(), ev, Pp * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_a.data(), SExp, (size_t)Pp * Pp * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_b.data(), SExp2, (size_t)Pp * Pp * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  double split_err = 0.0;
  for (int i = 0; i < P; ++i) for (int j = 0; j < P; ++j) {
    const double sc = std::sqrt(h_a[(size_t)i * Pp + i] * h_a[(size_t)j * Pp + j]);
    if (sc > 0) split_err = std::fmax(split_err, std::fabs(h_a[(size_t)i * Pp + j] - h_b[(size_t)i * Pp + j]) / sc);
  }
  const bool ok1 = std::isfinite(h_scal[0]) && h_scal[0] < 1e-6 && std::isfinite(h_scal[1]) && h_ev[P - 1] > 0 && split_err < 1e-6;
  // two-call eigensolver on a 512 x 512 Gram of the same O columns (blocked path)
  const int n2 = 512;
  double *G = dalloc<double>((size_t)n2 * n2), *X = dalloc<double>((size_t)N * n2), *G2 = dalloc<double>((size_t)n2 * n2);
  for (int rep = 0; rep < 4; ++rep)    // X = [O | O | O | O] (rank <= 37: a heavily rank-deficient matrix, like S)
    CU(cudaMemcpy2DAsync(X + rep * 128, (size_t)n2 * 8, O, (size_t)Pp * 8, (size_t)Pp * 8, N, cudaMemcpyDeviceToDevice, s));
  CK(vmcpde_syrk_tn(X, n2, G, n2, n2, N, 1.0 / N, 0.0, s));
  CK(vmcpde_sym_finalize(G, n2, 1.0, s));
  CU(cudaMemcpyAsync(G2, G, (size_t)n2 * n2 * 8, cudaMemcpyDeviceToDevice, s));
  size_t e2_b = 0; CK(vmcpde_eigh_workspace_bytes(n2, n2, &e2_b));
  void* ws2 = dalloc<char>(e2_b);
  double *ev2 = dalloc<double>(n2), *ZT = dalloc<double>((size_t)n2 * n2), *tau = dalloc<double>(n2), *VT2 = dalloc<double>((size_t)n2 * n2);
  CK(vmcpde_eigh_factor(G2, n2, n2, ev2, ZT, tau, ws2, e2_b, s));
  CK(vmcpde_eigh_backtransform(G2, tau, ZT, n2, n2, VT2, 0, 0, ws2, e2_b, s));
  std::vector<double> h_ev2(n2), h_G((size_t)n2 * n2), h_V((size_t)n2 * n2);
  CU(cudaMemcpyAsync(h_ev2.data(), ev2, n2 * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_G.data(), G, (size_t)n2 * n2 * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_V.data(), VT2, (size_t)n2 * n2 * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  double res = 0.0;   // || G v_k - ev_k v_k ||_max over the 8 largest eigenpairs
  for (int k = n2 - 8; k < n2; ++k)
    for (int i = 0; i < n2; ++i) {
      double acc = 0.0;
      for (int j = 0; j < n2; ++j) acc += h_G[(size_t)i * n2 + j] * h_V[(size_t)k * n2 + j];
      res = std::fmax(res, std::fabs(acc - h_ev2[k] * h_V[(size_t)k * n2 + i]));
    }
  const bool ok2 = res < 1e-11 * std::fabs(h_ev2[n2 - 1]) && h_ev2[n2 - 1] > 0;
  vmcpde_flow_destroy(flow);
  std::printf("%s residual %.3e tdvp_error %.6f ev_max %.6f split_err %.2e eig512_res %.2e\n", (ok1 && ok2) ? "CABI_SMOKE OK" : "CABI_SMOKE FAILED",
              h_scal[0], h_scal[1], h_ev[P - 1], split_err, res);
  return (ok1 && ok2) ? 0 : 1;
}
