"""End-to-end bring-up: TDVP RHS / Heun step through the reference-mirroring API vs the oracle."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, stepper, util, net, mpi_wrapper
from oracle import flow as oflow, tdvp as otdvp
def err(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a); b = np.asarray(b)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))

def build(d, depth, h, variant, latent, eqname, offset):
    smp = sampler.Sampler(dim=d, numChains=30, name=latent, mcmc_info={"offset": offset, "bound": 0.25})
    net.SingleBlock.different_add = (variant == "different_add"); net.SingleBlock.no_add = True
    vs = var_state.VarState(smp, d, 1, depth, network_args={"intmediate": (h,), "offset": offset, "latentSpaceName": latent, "dim": d})
    eq = evolutionEq.EvolutionEquation(dim=d, name=eqname)
    spec = oflow.FlowSpec(dim=d, depth=depth, hidden=(h,), latent=latent, variant=variant, offset=offset, inds_up=vs.net.inds_up, inds_down=vs.net.inds_down)
    assert spec.num_params == vs.numParameters, (spec.num_params, vs.numParameters)
    return smp, vs, eq, spec

def compare_rhs(d, depth, h, variant, latent, eqname, N, offset):
    smp, vs, eq, spec = build(d, depth, h, variant, latent, eqname, offset)
    theta = vs.get_parameters()
    th_np = theta.cpu().numpy()
    ost = oflow.OracleState(spec, th_np)
    OT = otdvp.OracleTDVP()
    upd_o, info_o = OT.rhs(ost, th_np, eqname, N)
    T = tdvp.TDVP()
    tm = util.Timings()
    upd, info = T(theta, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=tm)
    du = upd.cpu().numpy() - upd_o
    e = dict(S0=err(T.S0, OT.S0), SExp=err(T.SExp, OT.SExp), F=err(T.F0, OT.F0), ev=err(T.ev, OT.ev),
             upd_Snorm=float(du @ OT.S0 @ du / (upd_o @ OT.S0 @ upd_o)), res=float(T.solverResidual), res_o=OT.solverResidual,
             tdvp_err=abs(float(T.tdvp_error) - OT.tdvp_error), Emean=abs(float(T.ElocMean) - OT.ElocMean), Evar=abs(float(T.ElocVar) / OT.ElocVar - 1))
    for k in ("x1", "covar", "entropy", "x3", "x4", "x5", "x6", "max_grad", "integral_1sigma", "integral_0.5sigma", "integral_0.1sigma"):
        e["i_" + k] = err(info[k], info_o[k])
    big = np.abs(OT.ev / OT.ev[-1]) > 1e-6
    e["snr"] = float(np.abs(T.snr.cpu().numpy()[big] / OT.snr[big] - 1).max())
    print(d, depth, h, variant, latent, eqname, N, "P=%d" % vs.numParameters, {k: f"{v:.1e}" for k, v in e.items()}, flush=True)
    tm.print_timings()
    return vs, eq, T, ost, OT, spec

z2 = np.zeros(2)
compare_rhs(2, 4, 1, "no_add", "Gauss", "diffusion", 10000, z2)
compare_rhs(6, 4, 3, "different_add", "Gauss", "advection_hamiltonian_wDiss", 6000, np.array([1., 0, 0, 1, 0, 0]))
compare_rhs(8, 4, 4, "no_add", "Gauss", "diffusion", 5000, np.zeros(8))
compare_rhs(4, 3, 6, "no_add", "Gauss", "diffusion_anisotropic", 3000, np.zeros(4))

# Heun step vs oracle, chunked two-pass path (forced small chunks) and cholesky path
smp, vs, eq, spec = build(2, 4, 1, "no_add", "Gauss", "diffusion", z2)
theta0 = vs.get_parameters(); th_np = theta0.cpu().numpy()
ost = oflow.OracleState(spec, th_np); OT = otdvp.OracleTDVP()
f_o = lambda y, k: OT.rhs(ost, y, "diffusion", 4000, observables=False)[0]
y_o, dt_o = otdvp.heun_step(f_o, th_np, 1e-3, 1e-2, 1.3)
T = tdvp.TDVP(chunkSamples=1024)
st = stepper.FixedStepper(timeStep=1e-3, mode='Heun', maxStep=1e-2, increase_fac=1.3)
y, dt, info = st.step(0, T, theta0, evolutionEq=eq, psi=vs, nSamplesTDVP=4000, nSamplesObs=4000, normFunction=lambda v, S: v @ S @ v, timings=None)
print("heun (chunked two-pass): dt", dt, dt_o, "y err", err(y, y_o), "entropy", float(info["entropy"]), flush=True)
# adaptive heun
smp, vs, eq, spec = build(2, 4, 1, "no_add", "Gauss", "diffusion", z2)
ost = oflow.OracleState(spec, th_np); OT = otdvp.OracleTDVP()
f_o = lambda y, k: OT.rhs(ost, y, "diffusion", 3000, observables=False)[0]
y_o, rdt_o, ndt_o = otdvp.adaptive_heun_step(f_o, th_np, 1e-3, 1e-2, 1e-2, lambda v: v @ OT.SExp @ v)
T = tdvp.TDVP()
ah = stepper.AdaptiveHeun(timeStep=1e-3, tol=1e-2, maxStep=1e-2)
y, rdt, info = ah.step(0, T, theta0, evolutionEq=eq, psi=vs, nSamplesTDVP=3000, nSamplesObs=3000, normFunction=lambda v, S: v @ S @ v, timings=None)
print("adaptive heun: dt", rdt, rdt_o, "next", ah.dt, ndt_o, "y err", err(y, y_o), flush=True)
# cholesky path vs eigh path with a diagonal shift
smp, vs, eq, spec = build(6, 4, 3, "no_add", "Gauss", "diffusion", np.zeros(6))
theta0 = vs.get_parameters()
T1 = tdvp.TDVP(diagonalShift=1e-4); u1, _ = T1(theta0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=4000, nSamplesObs=4000, timings=None)
smp, vs, eq, spec = build(6, 4, 3, "no_add", "Gauss", "diffusion", np.zeros(6))
T2 = tdvp.TDVP(diagonalShift=1e-4, solver="cholesky"); u2, _ = T2(theta0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=4000, nSamplesObs=4000, timings=None)
print("cholesky vs eigh (shift 1e-4): update err", err(u2, u1.cpu().numpy()), "res", float(T1.solverResidual), float(T2.solverResidual), "tdvp_err", float(T1.tdvp_error), float(T2.tdvp_error), flush=True)

# C3-sized timing: d=6, depth 8, h=36, different_add, wDiss, N=2^18
import os
if os.environ.get("PROBE_C3", "1") == "1":
    off = np.array([1., 0, 0, 1, 0, 0])
    smp, vs, eq, spec = build(6, 8, 36, "different_add", "Gauss", "advection_hamiltonian_wDiss", off)
    print("C3 P =", vs.numParameters, flush=True)
    theta0 = vs.get_parameters()
    T = tdvp.TDVP()
    N = 2 ** 18
    for it in range(2):
        tm = util.Timings()
        torch.cuda.synchronize(); t0 = time.time()
        upd, info = T(theta0, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=tm)
        torch.cuda.synchronize(); print("C3 RHS sec", time.time() - t0, "res", float(T.solverResidual), "tdvp_err", float(T.tdvp_error), "entropy", float(info["entropy"]), "ev max", float(T.ev[-1]), "nbelow", int((T.ev / T.ev[-1] < 1e-11).sum()), flush=True)
        tm.print_timings()
