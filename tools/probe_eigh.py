"""GPU bring-up probe for vmcpde_eigh (run by hand under gpurun): accuracy against LAPACK and stage timings."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from vmc_pde_b200 import _lib
L = _lib.load(); dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:]] or [384, 600, 1000, 2053]


def run(S_np, label, check=True, reps=1):
    n = S_np.shape[0]; ld = L.vmcpde_padded_params(n)
    St = torch.zeros(ld, ld, device=dev, dtype=torch.float64); St[:n, :n] = torch.tensor(S_np, device=dev)
    ev = torch.zeros(ld, device=dev, dtype=torch.float64); VT = torch.zeros(ld, ld, device=dev, dtype=torch.float64)
    nb = C.c_size_t(0); _lib.check(L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(nb)))
    ws = torch.empty(nb.value, device=dev, dtype=torch.uint8)
    for rep in range(reps):
        A = St.clone()
        torch.cuda.synchronize(); t0 = time.time()
        _lib.check(L.vmcpde_eigh(_lib.ptr(A), n, ld, _lib.ptr(ev), _lib.ptr(VT), _lib.ptr(ws), nb.value, _lib.stream()))
        torch.cuda.synchronize(); dt = time.time() - t0
    V = VT[:n, :n].T
    S = St[:n, :n]
    nrm = float(S.abs().max()) * n ** 0.5 + 1e-300
    res = float((S @ V - V * ev[:n]).abs().max()) / nrm
    orth = float((V.T @ V - torch.eye(n, device=dev, dtype=torch.float64)).abs().max())
    msg = f"{label}: n={n} sec={dt:.4f} resid={res:.2e} orth={orth:.2e} sorted={bool((ev[1:n] >= ev[:n-1]).all())}"
    if check:
        ref = np.linalg.eigvalsh(S_np)
        msg += f" ev_err={np.abs(ev[:n].cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-300):.2e}"
    print(msg, flush=True)
    return ev, VT


rng = np.random.default_rng(0)
for n in sizes:
    if n <= 4200:
        A = rng.normal(size=(n, n)); run((A + A.T) / 2, "random sym")
    cs = 10.0 ** (-6.0 * np.arange(n) / n)
    Ot = torch.tensor(rng.normal(size=(2 * n, n)) * cs, device=dev); Sg = (Ot.T @ Ot / (2 * n)).cpu().numpy()
    run(Sg, "graded gram", check=n <= 4200, reps=2 if n > 4200 else 1)
    if n > 4200:
        St = torch.tensor(Sg, device=dev); torch.cuda.synchronize(); t0 = time.time()
        torch.linalg.eigh(St); torch.cuda.synchronize(); print("  torch eigh (cusolver) sec", time.time() - t0, flush=True)
        t0 = time.time(); torch.linalg.eigh(St); torch.cuda.synchronize(); print("  torch eigh (cusolver) 2nd sec", time.time() - t0, flush=True)
