import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, stepper, _kernels
np.random.seed(0)
off = np.zeros(8)
smp = sampler.Sampler(dim=8, numChains=30, name="Student_t", mcmc_info={"offset": off, "bound": 0.25})
vs = var_state.VarState(smp, 8, 1, 4, network_args={"intmediate": (4,), "offset": off, "latentSpaceName": "Student_t", "dim": 8})
eq = evolutionEq.EvolutionEquation(dim=8, name="diffusion")
T = tdvp.TDVP()
theta = vs.get_parameters()
dt = 1e-7
for k in range(400):
    dt = min(dt * 1.3, 1e-2)
    key_before = vs.sampler.key.copy()
    try:
        upd, info = T(theta, 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=10000, nSamplesObs=10000, timings=None)
    except SystemExit:
        print("NaN at step", k, "dt", dt)
        vs.sampler.key = key_before
        x, lp = vs.sample(10000)
        print("x finite", bool(torch.isfinite(x).all()), "max|x|", float(x.abs().max()), "lp finite", bool(torch.isfinite(lp).all()))
        E, O, lp2 = eq(vs, x, 0.0)
        print("E finite", bool(torch.isfinite(E).all()), "O finite", bool(torch.isfinite(O).all()), "lp2 finite", bool(torch.isfinite(lp2).all()))
        bad = ~torch.isfinite(E[0])
        print("n bad E", int(bad.sum()), "x of bad", x[0][bad][:3], "lp", lp2[0][bad][:3])
        print("S0 finite", bool(torch.isfinite(T.S0).all()), "F0 finite", bool(torch.isfinite(T.F0).all()), "ev", T.ev[:3], T.ev[-3:])
        print("theta dist/Ldiag", theta[:50])
        break
    theta = theta + dt * upd
    if k % 4 == 0:
        print(k, "dt %.2e ent %.4f dp %.4f err %.3e res %.2e |upd| %.3e argmax %d evmax %.3e evmin %.3e Fmax %.3e Emax %.3e" % (dt, float(info["entropy"]), float(vs.params["params"]["dist_params"][0]), float(T.tdvp_error), float(T.solverResidual), float(upd.abs().max()), int(upd.abs().argmax()), float(T.ev[-1]), float(T.ev[0]), float(T.F0.abs().max()), float(info["max_grad"])))
