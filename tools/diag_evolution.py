"""Diagnostic: TDVP evolution of the checked-in damped-oscillator physics against the stored particle run (prints a table).
usage (GPU box): python tools/diag_evolution.py [t_end] [N]"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmc_pde_b200 import sampler, var_state, evolutionEq, net, tdvp, stepper
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_wiener_T10.npz"))
t_end = float(sys.argv[1]) if len(sys.argv) > 1 else 12.0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
off = np.array([1.0, 0, 1, 0, 1, 0])
smp = sampler.Sampler(dim=6, numChains=30, name="Gauss", mcmc_info={"offset": off, "bound": 0.25})
net.SingleBlock.different_add = True
vs = var_state.VarState(smp, 6, 1, 4, network_args={"intmediate": (3,), "offset": off, "latentSpaceName": "Gauss", "dim": 6})
eq = evolutionEq.EvolutionEquation(dim=6, name="advection_hamiltonian_wDiss")
st = stepper.FixedStepper(timeStep=1e-4, mode='Heun', maxStep=1e-2, increase_fac=1.3)
T = tdvp.TDVP()
t, k, out = 0.0, 0, []
checks = [0.25, 0.5, 1, 1.5, 2, 3, 4, 5, 6, 8, 10, 12]
tw = g["times"]
np.set_printoptions(precision=3, suppress=True, linewidth=200)
import time
t0 = time.time()
while t < t_end + 1e-9:
    dp, dt, info = st.step(0, T, vs.get_parameters(), evolutionEq=eq, psi=vs, nSamplesTDVP=N, nSamplesObs=N, normFunction=lambda v, S: v @ S @ v, timings=None, integrals=False)
    vs.set_parameters(dp)
    k += 1
    if checks and t + dt >= checks[0]:
        checks.pop(0)
        tt = t + dt
        i = int(np.argmin(np.abs(tw - tt)))
        Cw = g["covar"][i]
        C = info["covar"].cpu().numpy()
        ent_w = 0.5 * np.linalg.slogdet(2 * np.pi * np.e * Cw)[1]
        rec = dict(t=tt, dx1=float(np.abs(info["x1"].cpu().numpy() - g["x1"][i]).max()), dcov=float(np.abs(np.diag(C) / np.diag(Cw) - 1).max()),
                   dxp=float(np.abs(C[0, 1] - Cw[0, 1])), ent=float(info["entropy"]), ent_w=float(ent_w),
                   i1=float(info["integral_1sigma"]), i1w=float(g["integral_1sigma"][i]), i05=float(info["integral_0.5sigma"]), i05w=float(g["integral_0.5sigma"][i]),
                   i01=float(info["integral_0.1sigma"]), tdvp_err=float(T.tdvp_error), res=float(T.solverResidual))
        out.append(rec)
        print(json.dumps(rec))
    t += dt
print("steps", k, "seconds", time.time() - t0)
