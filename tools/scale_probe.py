import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, ctypes as C
from vmc_pde_b200 import _lib
import test_gpu_kernels as tk
L = _lib.load()
rng = np.random.default_rng(7)
A = rng.normal(size=(500, 500)); S0 = (A + A.T) / 2
for sc in (1.0, 1e-6, 1e-12, 1e-140, 1e-170, 1e150):
    S = S0 * sc
    ev, V = tk._eigh(L, S)
    ref = np.linalg.eigvalsh(S); nrm = np.abs(ref).max()
    print(sc, "finite", np.isfinite(ev).all(), np.isfinite(V).all(), "ev_err", np.abs(ev - ref).max() / nrm,
          "res", np.abs(S @ V - V * ev).max() / nrm, "orth", np.abs(V.T @ V - np.eye(500)).max(), flush=True)
