"""Diagnostic: TDVP free diffusion with the Student-t latent (main.py mode 'diffusion', d = 8) against the reference's stored run.
usage (GPU box): python tools/diag_student.py [t_end]"""
import sys, os, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, stepper
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_diff8_student.npz"))
t_end = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
max_step = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-2
np.random.seed(0)
off = np.zeros(8)
smp = sampler.Sampler(dim=8, numChains=30, name="Student_t", mcmc_info={"offset": off, "bound": 0.25})
vs = var_state.VarState(smp, 8, 1, 4, network_args={"intmediate": (4,), "offset": off, "latentSpaceName": "Student_t", "dim": 8})
eq = evolutionEq.EvolutionEquation(dim=8, name="diffusion")
st = stepper.FixedStepper(timeStep=1e-7, mode='Heun', maxStep=max_step, increase_fac=1.3)
T = tdvp.TDVP()
t, k = 0.0, 0
checks = [0.01, 0.1, 0.5, 1, 2, 3, 5]
tw = g["times"]
t0 = time.time()
while t < t_end + 1e-9:
    dp, dt, info = st.step(0, T, vs.get_parameters(), evolutionEq=eq, psi=vs, nSamplesTDVP=10000, nSamplesObs=10000, normFunction=lambda v, S: v @ S @ v, timings=None, integrals=False)
    vs.set_parameters(dp)
    k += 1
    if checks and t + dt >= checks[0]:
        checks.pop(0)
        tt = t + dt
        i = int(np.argmin(np.abs(tw - tt)))
        print(json.dumps(dict(t=tt, t_ref=float(tw[i]), dp=float(vs.params["params"]["dist_params"][0]), dp_ref=float(g["dist_params"][i][0]),
                              ent=float(info["entropy"]), ent_ref=float(g["entropy"][i]), i1=float(info["integral_1sigma"]), i1_ref=float(g["integral_1sigma"][i]),
                              tdvp_err=float(T.tdvp_error), err_ref=float(g["tdvp_error"][i]), res=float(T.solverResidual))))
    t += dt
print("steps", k, "seconds", time.time() - t0, "P", vs.numParameters)
