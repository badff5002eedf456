// XLA-FFI custom-call handlers over the C-ABI of include/vmcpde.h, for the reference's JAX driver
// (north star: "thin jax.ffi custom calls").  Compiled ONLY when the XLA FFI headers exist
// (python -c "import jax; print(jax.ffi.include_dir())"); they are absent from this image, so this
// file is not part of libvmcpde.so here and has not been exercised.  build.py probes `import jax` on every build and
// compiles it into libvmcpde_xla.so when the headers exist (ten handlers: sample, local_terms, logp, moments1, center_force,
// gram, gram_split, finalize, eigh, solve_tail -- one per stage of TDVP.__call__, tdvp.py:96-164):
//     g++ -shared -fPIC -std=c++17 -I$(jax.ffi.include_dir) -Iinclude xla_ffi_shim.cc -L. -lvmcpde -o libvmcpde_xla.so
// Registration on the Python side is shown in INTEGRATION.md.
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "xla/ffi/api/ffi.h"
#include "../../include/vmcpde.h"

namespace ffi = xla::ffi;

namespace {

// vmcpde_flow handles are cached per architecture (the attrs of every call carry the architecture)
struct FlowKey {
  std::vector<int32_t> ints;
  std::vector<double> offset;
  bool operator<(const FlowKey& o) const { return ints != o.ints ? ints < o.ints : offset < o.offset; }
};
std::mutex g_mu;
std::map<FlowKey, vmcpde_flow*> g_flows;

vmcpde_flow* get_flow(int32_t dim, int32_t depth, int32_t hidden, int32_t variant, int32_t latent,
                      ffi::Span<const int32_t> ind_up, ffi::Span<const int32_t> ind_down, ffi::Span<const double> offset) {
  FlowKey k;
  k.ints = {dim, depth, hidden, variant, latent};
  k.ints.insert(k.ints.end(), ind_up.begin(), ind_up.end());
  k.ints.insert(k.ints.end(), ind_down.begin(), ind_down.end());
  k.offset.assign(offset.begin(), offset.end());
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_flows.find(k);
  if (it != g_flows.end()) return it->second;
  vmcpde_flow_config cfg{dim, depth, 1, hidden, variant, latent, ind_up.begin(), ind_down.begin(), offset.begin(), nullptr};
  vmcpde_flow* f = nullptr;
  if (vmcpde_flow_create(&cfg, &f) != 0) return nullptr;
  g_flows[k] = f;
  return f;
}

ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, vmcpde_last_error());
}

#define FLOW_ATTRS(b)                                                                                          \
  b.Attr<int32_t>("dim").Attr<int32_t>("depth").Attr<int32_t>("hidden").Attr<int32_t>("variant").Attr<int32_t>("latent") \
   .Attr<ffi::Span<const int32_t>>("ind_up").Attr<ffi::Span<const int32_t>>("ind_down").Attr<ffi::Span<const double>>("offset")

// (1) sampler: theta[P], key[2] (uint32) -> x[n,d], logp[n]
ffi::Error SampleImpl(cudaStream_t stream, int32_t dim, int32_t depth, int32_t hidden, int32_t variant, int32_t latent,
                      ffi::Span<const int32_t> ind_up, ffi::Span<const int32_t> ind_down, ffi::Span<const double> offset,
                      int64_t first, int64_t n_total, uint32_t key0, uint32_t key1, ffi::Buffer<ffi::F64> theta,
                      ffi::Buffer<ffi::F64> chi2, ffi::ResultBuffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> logp) {
  vmcpde_flow* f = get_flow(dim, depth, hidden, variant, latent, ind_up, ind_down, offset);
  if (!f) return ffi::Error(ffi::ErrorCode::kInvalidArgument, vmcpde_last_error());
  const int64_t n = logp->element_count();
  const double* c2 = chi2.element_count() > 0 ? chi2.typed_data() : nullptr;
  return status(vmcpde_sample(f, theta.typed_data(), key0, key1, first, n, n_total, c2, x->typed_data(), logp->typed_data(),
                              nullptr, stream));
}

// (2) fused local terms: theta[P], x[n,d] -> eloc[n], logp[n], O[n, ldo]
ffi::Error LocalTermsImpl(cudaStream_t stream, int32_t dim, int32_t depth, int32_t hidden, int32_t variant, int32_t latent,
                          ffi::Span<const int32_t> ind_up, ffi::Span<const int32_t> ind_down, ffi::Span<const double> offset,
                          int32_t mode, ffi::Span<const double> eq /* D, mu, m, omega, lam, T, gamma, t */,
                          ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> tangents,
                          ffi::ResultBuffer<ffi::F64> eloc, ffi::ResultBuffer<ffi::F64> logp, ffi::ResultBuffer<ffi::F64> O) {
  vmcpde_flow* f = get_flow(dim, depth, hidden, variant, latent, ind_up, ind_down, offset);
  if (!f) return ffi::Error(ffi::ErrorCode::kInvalidArgument, vmcpde_last_error());
  vmcpde_equation e{mode, eq[0], eq[1], eq[2], eq[3], eq[4], eq[5], eq[6], eq[7],
                    tangents.element_count() > 0 ? tangents.typed_data() : nullptr};
  const int64_t n = eloc->element_count();
  const int64_t ldo = O->dimensions().back();
  return status(vmcpde_local_terms(f, theta.typed_data(), x.typed_data(), n, &e, eloc->typed_data(), logp->typed_data(),
                                   nullptr, nullptr, O->typed_data(), ldo, stream));
}

// (3) weighted Gram accumulation: O[n, ldo], w[n] (or empty), S[Pp, Pp] (aliased in/out)
ffi::Error GramImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> O, ffi::Buffer<ffi::F64> w, ffi::Buffer<ffi::F64> S_in,
                    ffi::ResultBuffer<ffi::F64> S) {
  const int64_t n = O.dimensions()[0], ldo = O.dimensions()[1];
  const int32_t Pp = (int32_t)S->dimensions()[0];
  if (S_in.typed_data() != S->typed_data())
    cudaMemcpyAsync(S->typed_data(), S_in.typed_data(), sizeof(double) * Pp * Pp, cudaMemcpyDeviceToDevice, stream);
  const double* wp[1] = {w.element_count() > 0 ? w.typed_data() : nullptr};
  double* sp[1] = {S->typed_data()};
  return status(vmcpde_gram(O.typed_data(), n, ldo, Pp, 1, wp, sp, stream));
}

// (4) eigendecomposition: S[n, ld] -> ev[ld], VT[ld, ld]; scratch is an XLA-allocated result buffer
ffi::Error EighImpl(cudaStream_t stream, int32_t n, ffi::Buffer<ffi::F64> S, ffi::ResultBuffer<ffi::F64> ev,
                    ffi::ResultBuffer<ffi::F64> VT, ffi::ResultBuffer<ffi::F64> work, ffi::ResultBuffer<ffi::U8> scratch) {
  const int32_t ld = (int32_t)S.dimensions()[1];
  cudaMemcpyAsync(work->typed_data(), S.typed_data(), sizeof(double) * ld * ld, cudaMemcpyDeviceToDevice, stream);
  cudaMemsetAsync(VT->typed_data(), 0, sizeof(double) * ld * ld, stream);
  return status(vmcpde_eigh(work->typed_data(), n, ld, ev->typed_data(), VT->typed_data(), scratch->typed_data(),
                            scratch->element_count(), stream));
}

// (5) first moments (tdvp.py:37-41): eloc[n], logp[n], O[n, ldo] -> sums[4 + ldo]; scratch from XLA
ffi::Error Moments1Impl(cudaStream_t stream, ffi::Buffer<ffi::F64> eloc, ffi::Buffer<ffi::F64> logp, ffi::Buffer<ffi::F64> O,
                        ffi::ResultBuffer<ffi::F64> sums, ffi::ResultBuffer<ffi::U8> scratch) {
  const int64_t n = O.dimensions()[0], ldo = O.dimensions()[1];
  cudaMemsetAsync(sums->typed_data(), 0, sizeof(double) * (4 + ldo), stream);
  return status(vmcpde_moments1(eloc.typed_data(), logp.typed_data(), O.typed_data(), n, ldo, sums->typed_data(),
                                scratch->typed_data(), scratch->element_count(), stream));
}

// (6) centring + force (tdvp.py:40-45): O[n, ldo] (aliased in/out), meanO[ldo], eloc[n], logp[n], attr meanE
//     -> O centred, dE[n], wE[n], wLp[n], Fsum[ldo], var_sum[8]
ffi::Error CenterForceImpl(cudaStream_t stream, double meanE, ffi::Buffer<ffi::F64> O_in, ffi::Buffer<ffi::F64> meanO,
                           ffi::Buffer<ffi::F64> eloc, ffi::Buffer<ffi::F64> logp, ffi::ResultBuffer<ffi::F64> O,
                           ffi::ResultBuffer<ffi::F64> dE, ffi::ResultBuffer<ffi::F64> wE, ffi::ResultBuffer<ffi::F64> wLp,
                           ffi::ResultBuffer<ffi::F64> Fsum, ffi::ResultBuffer<ffi::F64> var_sum, ffi::ResultBuffer<ffi::U8> scratch) {
  const int64_t n = O_in.dimensions()[0], ldo = O_in.dimensions()[1];
  if (O_in.typed_data() != O->typed_data())
    cudaMemcpyAsync(O->typed_data(), O_in.typed_data(), sizeof(double) * n * ldo, cudaMemcpyDeviceToDevice, stream);
  cudaMemsetAsync(Fsum->typed_data(), 0, sizeof(double) * ldo, stream);
  cudaMemsetAsync(var_sum->typed_data(), 0, sizeof(double) * var_sum->element_count(), stream);
  return status(vmcpde_center_force(O->typed_data(), n, ldo, meanO.typed_data(), eloc.typed_data(), logp.typed_data(), meanE,
                                    dE->typed_data(), wE->typed_data(), wLp->typed_data(), Fsum->typed_data(), var_sum->typed_data(),
                                    scratch->typed_data(), scratch->element_count(), stream));
}

// (7) split-precision Gram on tcgen05 (SExp, SNR covariance): like (3) with a bf16-slice workspace from XLA
ffi::Error GramSplitImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> O, ffi::Buffer<ffi::F64> w, ffi::Buffer<ffi::F64> S_in,
                         ffi::ResultBuffer<ffi::F64> S, ffi::ResultBuffer<ffi::U8> scratch) {
  const int64_t n = O.dimensions()[0], ldo = O.dimensions()[1];
  const int32_t Pp = (int32_t)S->dimensions()[0];
  if (S_in.typed_data() != S->typed_data())
    cudaMemcpyAsync(S->typed_data(), S_in.typed_data(), sizeof(double) * Pp * Pp, cudaMemcpyDeviceToDevice, stream);
  return status(vmcpde_gram_split(O.typed_data(), n, ldo, Pp, w.element_count() > 0 ? w.typed_data() : nullptr, S->typed_data(),
                                  scratch->typed_data(), scratch->element_count(), stream));
}

// (8) S <- scale * upper(S) mirrored (+ multiplicative diagonal shift, tdvp.py:50-51): S (aliased) -> S, S_shifted
ffi::Error FinalizeImpl(cudaStream_t stream, double scale, double shift, int32_t P, ffi::Buffer<ffi::F64> S_in,
                        ffi::ResultBuffer<ffi::F64> S, ffi::ResultBuffer<ffi::F64> S_shifted) {
  const int32_t Pp = (int32_t)S->dimensions()[0];
  if (S_in.typed_data() != S->typed_data())
    cudaMemcpyAsync(S->typed_data(), S_in.typed_data(), sizeof(double) * Pp * Pp, cudaMemcpyDeviceToDevice, stream);
  if (int rc = vmcpde_sym_finalize(S->typed_data(), Pp, scale, stream)) return status(rc);
  return status(vmcpde_diag_shift(S->typed_data(), S_shifted->typed_data(), Pp, P, shift, stream));
}

// (9) everything after eigh (tdvp.py:66-94): ev[ld], VT[ld, ld], F[ld], S, S0, CEO (or empty) -> VtF, rhoVar, snr, invEv,
//     update (all [ld]), scalars[2] = (solver residual, tdvp_error)
ffi::Error SolveTailImpl(cudaStream_t stream, int32_t n, double n_glob, double svdTol, double snrTol, int32_t useSNR, double meanE2,
                         ffi::Buffer<ffi::F64> ev, ffi::Buffer<ffi::F64> VT, ffi::Buffer<ffi::F64> F, ffi::Buffer<ffi::F64> S,
                         ffi::Buffer<ffi::F64> S0, ffi::Buffer<ffi::F64> CEO, ffi::ResultBuffer<ffi::F64> VtF,
                         ffi::ResultBuffer<ffi::F64> rhoVar, ffi::ResultBuffer<ffi::F64> snr, ffi::ResultBuffer<ffi::F64> invEv,
                         ffi::ResultBuffer<ffi::F64> update, ffi::ResultBuffer<ffi::F64> scalars, ffi::ResultBuffer<ffi::U8> scratch) {
  const int32_t ld = (int32_t)VT.dimensions()[1];
  const bool ceo = CEO.element_count() > 0;
  return status(vmcpde_solve_tail(ev.typed_data(), VT.typed_data(), n, ld, F.typed_data(), S.typed_data(), S0.typed_data(),
                                  ceo ? CEO.typed_data() : nullptr, n_glob, svdTol, snrTol, useSNR, meanE2, VtF->typed_data(),
                                  ceo ? rhoVar->typed_data() : nullptr, ceo ? snr->typed_data() : nullptr, invEv->typed_data(),
                                  update->typed_data(), scalars->typed_data(), scratch->typed_data(), scratch->element_count(), stream));
}

// (10) log p(x) alone (VarState.__call__(mode="eval"), var_state.py:38-43; the ball integrals of tdvp.py:152-162)
ffi::Error LogpImpl(cudaStream_t stream, int32_t dim, int32_t depth, int32_t hidden, int32_t variant, int32_t latent,
                    ffi::Span<const int32_t> ind_up, ffi::Span<const int32_t> ind_down, ffi::Span<const double> offset,
                    ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> x, ffi::ResultBuffer<ffi::F64> logp) {
  vmcpde_flow* f = get_flow(dim, depth, hidden, variant, latent, ind_up, ind_down, offset);
  if (!f) return ffi::Error(ffi::ErrorCode::kInvalidArgument, vmcpde_last_error());
  return status(vmcpde_logp(f, theta.typed_data(), x.typed_data(), logp->element_count(), logp->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_sample, SampleImpl,
                              FLOW_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>())
                                  .Attr<int64_t>("first").Attr<int64_t>("n_total").Attr<uint32_t>("key0").Attr<uint32_t>("key1")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_local_terms, LocalTermsImpl,
                              FLOW_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>())
                                  .Attr<int32_t>("mode").Attr<ffi::Span<const double>>("eq")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_gram, GramImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_eigh, EighImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int32_t>("n")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_moments1, Moments1Impl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_center_force, CenterForceImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<double>("meanE")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_gram_split, GramSplitImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_finalize, FinalizeImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<double>("scale").Attr<double>("shift").Attr<int32_t>("P")
                                  .Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_solve_tail, SolveTailImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int32_t>("n").Attr<double>("n_glob")
                                  .Attr<double>("svdTol").Attr<double>("snrTol").Attr<int32_t>("useSNR").Attr<double>("meanE2")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(vmcpde_xla_logp, LogpImpl,
                              FLOW_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>())
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
