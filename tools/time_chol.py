"""Times vmcpde_chol_solve at the bench size (run by hand under gpurun)."""
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from vmc_pde_b200 import _lib
L = _lib.load(); dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8187
ld = L.vmcpde_padded_params(n)
g = torch.Generator(device=dev); g.manual_seed(0)
O = torch.randn(2 * n, n, device=dev, dtype=torch.float64, generator=g)
Sn = O.T @ O / (2 * n) + 1e-3 * torch.eye(n, device=dev, dtype=torch.float64)
F = torch.randn(n, device=dev, dtype=torch.float64, generator=g)
S0 = torch.zeros(ld, ld, device=dev, dtype=torch.float64); S0[:n, :n] = Sn
x = torch.zeros(n, device=dev, dtype=torch.float64); info = torch.zeros(1, device=dev, dtype=torch.int32)
for rep in range(3):
    S = S0.clone(); info.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    _lib.check(L.vmcpde_chol_solve(_lib.ptr(S), n, ld, _lib.ptr(F), _lib.ptr(x), _lib.ptr(info), _lib.stream()))
    e1.record(); torch.cuda.synchronize()
    r = float((Sn @ x - F).norm() / F.norm())
    print(f"chol_solve n={n}: {e0.elapsed_time(e1):.2f} ms, info={int(info)}, residual {r:.2e}", flush=True)
