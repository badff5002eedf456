"""First GPU bring-up probe (run by hand under gpurun; the pytest suite supersedes it)."""
import ctypes as C, sys, time, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from vmc_pde_b200 import _lib, _capi
from oracle import flow, tdvp, threefry
L = _lib.load()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
res = {}
def err(a, b): return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))

tf = C.c_double(0); _lib.check(L.vmcpde_dmma_peak(C.byref(tf))); res["dmma_peak_tflops"] = tf.value
print("DMMA peak", tf.value, flush=True)

rng = np.random.default_rng(0)
def case(d, depth, h, variant, latent, eqname, n):
    ups, downs, _ = flow.make_index_splits(d, depth, 1)
    off = rng.normal(size=d) * 0.3
    spec = flow.FlowSpec(dim=d, depth=depth, hidden=(h,), latent=latent, variant=variant, offset=off, inds_up=ups, inds_down=downs)
    th = flow.init_params(spec, 1) + 0.05 * rng.normal(size=spec.num_params)
    st = flow.OracleState(spec, th)
    cfg, keep = _capi.make_flow_config(d, depth, (h,), variant, latent, ups, downs, off)
    fh = C.c_void_p(); _lib.check(L.vmcpde_flow_create(C.byref(cfg), C.byref(fh)))
    P = L.vmcpde_flow_num_params(fh); assert P == spec.num_params
    Pp = L.vmcpde_padded_params(P)
    tht = torch.tensor(th, device=dev)
    # sampler
    x = torch.empty(n, d, device=dev, dtype=torch.float64); lp = torch.empty(n, device=dev, dtype=torch.float64); z = torch.empty_like(x)
    chi2 = None
    if latent == "Student_t":
        chi2_np = np.random.default_rng(5).chisquare(float(np.exp(th[spec.slices()[0]["dist_params"][0]]) + 1), size=n)
        chi2 = torch.tensor(chi2_np, device=dev); st.chi2 = lambda nu, m: chi2_np
    key = threefry.split(st.key, 2)[1]
    _lib.check(L.vmcpde_sample(fh, _lib.ptr(tht), int(key[0]), int(key[1]), 0, n, n, _lib.ptr(chi2), _lib.ptr(x), _lib.ptr(lp), _lib.ptr(z), _lib.stream()))
    xo, lpo, zo = st.sample(n)
    e = {"z": err(z.cpu(), zo), "x": err(x.cpu(), xo), "lp_s": err(lp.cpu(), lpo)}
    # local terms on the oracle's samples
    xin = torch.tensor(xo.numpy(), device=dev)
    A = None
    if eqname == "diffusion_anisotropic":
        A = torch.tensor(tdvp.random_D_factor(d), device=dev)
    eq = _capi.make_equation(eqname, dict(tdvp.EQ_PARAMS.get(eqname, {})), 0.3, A.data_ptr() if A is not None else None)
    nrow = (n + 15) // 16 * 16
    E = torch.empty(n, device=dev, dtype=torch.float64); lp2 = torch.empty_like(E); g = torch.empty(n, d, device=dev, dtype=torch.float64); lap = torch.empty_like(E)
    O = torch.zeros(nrow, Pp, device=dev, dtype=torch.float64)
    t0 = time.time()
    _lib.check(L.vmcpde_local_terms(fh, _lib.ptr(tht), _lib.ptr(xin), n, C.byref(eq), _lib.ptr(E), _lib.ptr(lp2), _lib.ptr(g), _lib.ptr(lap), _lib.ptr(O), Pp, _lib.stream()))
    torch.cuda.synchronize()
    Eo, Oo, lpo2, go = tdvp.local_terms(st, xo, eqname, 0.3)
    e.update(E=err(E.cpu(), Eo), O=err(O[:n, :P].cpu(), Oo), lp=err(lp2.cpu(), lpo2), pad=float(O[:, P:].abs().max()) if Pp > P else 0.0)
    if eqname != "diffusion_anisotropic": e["g"] = err(g.cpu(), go)
    # hessian + logp
    H = torch.empty(min(n, 64), d, d, device=dev, dtype=torch.float64)
    _lib.check(L.vmcpde_hessian(fh, _lib.ptr(tht), _lib.ptr(xin), min(n, 64), _lib.ptr(H), _lib.stream()))
    e["H"] = err(H.cpu(), st.hessian(xo[:64]))
    # moments / centre / gram
    sums = torch.zeros(4 + Pp, device=dev, dtype=torch.float64)
    mws = torch.empty(((n + 511) // 512 + 1) * Pp, device=O.device, dtype=torch.float64)
    _lib.check(L.vmcpde_moments1(_lib.ptr(E), _lib.ptr(lp2), _lib.ptr(O), n, Pp, _lib.ptr(sums), _lib.ptr(mws), mws.numel() * 8, _lib.stream()))
    T = tdvp.OracleTDVP(); T.solve(Eo.numpy(), Oo.numpy(), lpo2.numpy())
    e["meanE"] = abs(float(sums[0]) / n - T.ElocMean) / (abs(T.ElocMean) + 1e-300)
    e["meanO"] = err((sums[4:4 + P] / n).cpu(), T.gradMean)
    meanO = (sums[4:] / n).contiguous()
    dE = torch.zeros(nrow, device=dev, dtype=torch.float64); wE = torch.zeros_like(dE); wLp = torch.zeros_like(dE)
    F = torch.zeros(Pp, device=dev, dtype=torch.float64); var = torch.zeros(1, device=dev, dtype=torch.float64)
    _lib.check(L.vmcpde_center_force(_lib.ptr(O), n, Pp, _lib.ptr(meanO), _lib.ptr(E), _lib.ptr(lp2), float(sums[0]) / n, _lib.ptr(dE), _lib.ptr(wE), _lib.ptr(wLp), _lib.ptr(F), _lib.ptr(var), _lib.ptr(mws), mws.numel() * 8, _lib.stream()))
    e["F"] = err((F[:P] / n).cpu(), T.F0); e["var"] = abs(float(var) / n - T.ElocVar) / T.ElocVar
    S = [torch.zeros(Pp, Pp, device=dev, dtype=torch.float64) for _ in range(3)]
    _lib.check(L.vmcpde_gram(_lib.ptr(O), nrow, Pp, Pp, 3, _lib.ptr_array([None, wLp, wE]), _lib.ptr_array(S), _lib.stream()))
    for s_ in S: _lib.check(L.vmcpde_sym_finalize(_lib.ptr(s_), Pp, 1.0 / n, _lib.stream()))
    torch.cuda.synchronize()
    e["S0"] = err(S[0][:P, :P].cpu(), T.S0); e["SExp"] = err(S[1][:P, :P].cpu(), T.SExp)
    dO = Oo.numpy() - T.gradMean; dEo = Eo.numpy() - T.ElocMean
    Ceo = (dO * (dEo ** 2)[:, None]).T @ dO / n
    e["CEO"] = err(S[2][:P, :P].cpu(), Ceo)
    e["sym"] = float((S[0] - S[0].T).abs().max())
    L.vmcpde_flow_destroy(fh)
    print(d, depth, h, variant, latent, eqname, n, {k: f"{v:.1e}" for k, v in e.items()}, flush=True)
    return e

res["cases"] = []
for args in [(2, 4, 1, "no_add", "Gauss", "diffusion", 1000),
             (6, 3, 5, "different_add", "Gauss", "advection_hamiltonian_wDiss", 777),
             (8, 4, 4, "no_add", "Student_t", "diffusion", 2048),
             (4, 2, 3, "add_s", "Gauss", "diffusion_anisotropic", 515),
             (10, 2, 20, "no_add", "Gauss", "diffusion_drift", 300),
             (3, 2, 3, "jac_eq_1", "Gauss", "advection_hamiltonian", 100)]:
    try:
        res["cases"].append(case(*args))
    except Exception as ex:
        print("FAILED", args, repr(ex), flush=True); res["cases"].append({"fail": repr(ex)})

# Gram throughput on synthetic O
for (n, Pp, nm) in [(16384, 2048, 1), (16384, 8192, 1), (16384, 8192, 3), (65536, 8192, 3)]:
    O = torch.randn(n, Pp, device=dev, dtype=torch.float64)
    w = [None] + [torch.rand(n, device=dev, dtype=torch.float64) for _ in range(nm - 1)]
    S = [torch.zeros(Pp, Pp, device=dev, dtype=torch.float64) for _ in range(nm)]
    for it in range(2):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(L.vmcpde_gram(_lib.ptr(O), n, Pp, Pp, nm, _lib.ptr_array(w), _lib.ptr_array(S), _lib.stream())); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); fl = nm * n * Pp * (Pp + 1.0)
    ref = (O[:, :256].T @ O) ; chk = err((S[0][:256, :] / 2).cpu() if False else torch.triu(S[0] / 2)[:256].cpu(), torch.triu(ref)[:256].cpu())
    print(f"gram n={n} Pp={Pp} mats={nm}: {ms:.2f} ms  {fl / ms * 1e-9:.2f} TFLOP/s (SYRK flops)  err={chk:.1e}", flush=True)
    res.setdefault("gram", []).append({"n": n, "Pp": Pp, "mats": nm, "ms": ms, "tflops": fl / ms * 1e-9, "err": chk})
    # cuBLAS yardstick
    if nm == 1:
        for it in range(2):
            e0.record(); R = O.T @ O; e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1); print(f"   torch/cuBLAS DGEMM full: {ms2:.2f} ms {2.0*n*Pp*Pp/ms2*1e-9:.2f} TFLOP/s", flush=True)
        res["gram"][-1]["cublas_ms"] = ms2
    del O, S, w
json.dump(res, open("/root/repo/gpurun_out/probe1.json", "w"), indent=1)
