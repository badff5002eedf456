"""Cholesky-solver right-hand side on the bench workload (run by hand under gpurun): does the shifted S factorise?"""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from vmc_pde_b200 import tdvp
vs, eq, T, st = bench.build_ours()
N = 2 ** 16
for shift in (1e-4, 1e-2):
    Tc = tdvp.TDVP(diagonalShift=shift, solver="cholesky")
    try:
        upd, info = Tc(vs.get_parameters(), 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=N, nSamplesObs=N, timings=None)
        print("shift", shift, "ok residual", float(Tc.solverResidual), "tdvp_error", float(Tc.tdvp_error), flush=True)
    except RuntimeError as e:
        print("shift", shift, "FAILED:", e, flush=True)
        S = Tc.S.clone()
        d = torch.diagonal(S)
        print("  diag min/max", float(d.min()), float(d.max()), "asym", float((S - S.T).abs().max()), flush=True)
        ev = torch.linalg.eigvalsh(S)
        print("  eig min/max (torch)", float(ev[0]), float(ev[-1]), flush=True)
        Dm = torch.diag(1.0 / d.sqrt())
        evc = torch.linalg.eigvalsh(Dm @ S @ Dm)
        print("  scaled (correlation form) eig min/max", float(evc[0]), float(evc[-1]), flush=True)
