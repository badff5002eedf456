"""One TDVP right-hand side of the bench workload (C3) without warm-up -- the command profiled with ncu.
usage: python tools/profile_step.py [n_samples] [fp64|split]"""
import sys, json, time
import torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vmc_pde_b200 import _kernels
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.C3["n_samples"]
vs, eq, T, st = bench.build_ours("C3", gramPrecision=sys.argv[2] if len(sys.argv) > 2 else "fp64")
torch.cuda.synchronize(); t0 = time.time()
upd, info = T(vs.get_parameters(), 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=n, nSamplesObs=n, timings=None)
torch.cuda.synchronize()
print(json.dumps({"n": n, "sec": time.time() - t0, "launches": _kernels.launches, "residual": float(T.solverResidual)}))
