"""GPU bring-up probe for eigh / solve tail / cholesky / gemm_tn (run by hand under gpurun)."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from vmc_pde_b200 import _lib
from oracle import tdvp as otdvp
L = _lib.load(); dev = torch.device("cuda:0")
def err(a, b): return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))
def pad(n): return L.vmcpde_padded_params(n)
rng = np.random.default_rng(0)

def run_eigh(S_np, label, check=True):
    n = S_np.shape[0]; ld = pad(n)
    S = torch.zeros(ld, ld, device=dev, dtype=torch.float64); S[:n, :n] = torch.tensor(S_np, device=dev)
    ev = torch.zeros(ld, device=dev, dtype=torch.float64); VT = torch.zeros(ld, ld, device=dev, dtype=torch.float64)
    nb = C.c_size_t(0); _lib.check(L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(nb)))
    ws = torch.empty(nb.value, device=dev, dtype=torch.uint8)
    A = S.clone()
    torch.cuda.synchronize(); t0 = time.time()
    _lib.check(L.vmcpde_eigh(_lib.ptr(A), n, ld, _lib.ptr(ev), _lib.ptr(VT), _lib.ptr(ws), nb.value, _lib.stream()))
    torch.cuda.synchronize(); t1 = time.time()
    out = {"n": n, "sec": t1 - t0}
    if check:
        evn = ev[:n].cpu().numpy(); V = VT[:n, :n].cpu().numpy().T
        ref = np.linalg.eigvalsh(S_np); nrm = abs(ref).max()
        out.update(ev=float(np.abs(evn - ref).max() / nrm), resid=float(np.abs(S_np @ V - V * evn).max() / nrm),
                   orth=float(np.abs(V.T @ V - np.eye(n)).max()), sorted=bool(np.all(np.diff(evn) >= 0)))
    print(label, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in out.items()}, flush=True)
    return ev, VT

for n in (1, 2, 3, 37, 130, 300, 1000):
    A = rng.normal(size=(n, n)); run_eigh((A + A.T) / 2, "random sym")
A = rng.normal(size=(3000, 200)) @ rng.normal(size=(200, 600)); run_eigh(A.T @ A / 3000, "rank-deficient gram 600")
cs = 10.0 ** (-6.0 * np.arange(2053) / 2053); A = rng.normal(size=(6000, 2053)) * cs; run_eigh(A.T @ A / 6000, "graded gram 2053")
n = 8187
At = torch.randn(16384, n, device=dev, dtype=torch.float64) * torch.tensor(10.0 ** (-6.0 * np.arange(n) / n), device=dev)
St = (At.T @ At / 16384); S_np = St.cpu().numpy(); del At
ev, VT = run_eigh(S_np, "graded gram 8187", check=False)
ev, VT = run_eigh(S_np, "graded gram 8187 (2nd)", check=False)
V = VT[:n, :n].T; R = St @ V - V * ev[:n]
print("  8187: resid", float(R.abs().max() / ev[:n].abs().max()), "orth", float((V.T @ V - torch.eye(n, device=dev, dtype=torch.float64)).abs().max()), flush=True)
t0 = time.time(); evt = torch.linalg.eigvalsh(St); torch.cuda.synchronize(); print("  torch eigvalsh(cusolver) sec", time.time() - t0, "ev err", float((evt - ev[:n]).abs().max() / evt.abs().max()))
t0 = time.time(); evt, Vt = torch.linalg.eigh(St); torch.cuda.synchronize(); print("  torch eigh(cusolver) sec", time.time() - t0, flush=True)

# gemm_tn
K, M, N = 512, 256, 384
X = torch.randn(K, M, device=dev, dtype=torch.float64); Y = torch.randn(K, N, device=dev, dtype=torch.float64); O = torch.randn(M, N, device=dev, dtype=torch.float64)
ref = 0.5 * X.T @ Y + 2.0 * O
_lib.check(L.vmcpde_gemm_tn(_lib.ptr(X), M, _lib.ptr(Y), N, _lib.ptr(O), N, M, N, K, 0.5, 2.0, _lib.stream()))
print("gemm_tn err", err(O.cpu(), ref.cpu()), flush=True)

# solve tail vs oracle
for (n, Ns) in ((37, 2000), (300, 4000)):
    ld = pad(n)
    O_ = rng.normal(size=(Ns, n)) * 10.0 ** (-3.0 * np.arange(n) / n); O_[:, n // 2:] = O_[:, : n - n // 2] @ rng.normal(size=(n - n // 2, n - n // 2)) * 1e-1  # rank deficient
    E = rng.normal(size=Ns) + 0.3; lp = rng.normal(size=Ns)
    T = otdvp.OracleTDVP(); upd = T.solve(E, O_, lp)
    dO = O_ - O_.mean(0); dE = E - E.mean(); CEO = (dO * (dE ** 2)[:, None]).T @ dO / Ns
    def P(a):
        t = torch.zeros(ld, ld, device=dev, dtype=torch.float64); t[:n, :n] = torch.tensor(a, device=dev); return t
    S = P(T.S); S0 = P(T.S0); Cg = P(CEO); F = torch.zeros(ld, device=dev, dtype=torch.float64); F[:n] = torch.tensor(T.F0, device=dev)
    ev = torch.zeros(ld, device=dev, dtype=torch.float64); VT = torch.zeros(ld, ld, device=dev, dtype=torch.float64)
    nb = C.c_size_t(0); L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(nb)); nb2 = C.c_size_t(0); L.vmcpde_solve_tail_workspace_bytes(n, ld, C.byref(nb2))
    ws = torch.empty(max(nb.value, nb2.value), device=dev, dtype=torch.uint8)
    A = S.clone()
    _lib.check(L.vmcpde_eigh(_lib.ptr(A), n, ld, _lib.ptr(ev), _lib.ptr(VT), _lib.ptr(ws), ws.numel(), _lib.stream()))
    outs = [torch.zeros(ld, device=dev, dtype=torch.float64) for _ in range(5)]; sc = torch.zeros(2, device=dev, dtype=torch.float64)
    _lib.check(L.vmcpde_solve_tail(_lib.ptr(ev), _lib.ptr(VT), n, ld, _lib.ptr(F), _lib.ptr(S), _lib.ptr(S0), _lib.ptr(Cg), float(Ns), 1e-11, 2.0, 0, float(np.mean(E ** 2)),
                                   *[_lib.ptr(o) for o in outs], _lib.ptr(sc), _lib.ptr(ws), ws.numel(), _lib.stream()))
    VtF, rhoVar, snr, invEv, update = [o[:n].cpu().numpy() for o in outs]
    u = update; Sn = T.S
    print(f"solve_tail n={n}: ev {err(ev[:n].cpu(), T.ev):.1e} update-vs-oracle(S-norm) {float((u - upd) @ Sn @ (u - upd) / (upd @ Sn @ upd)):.1e} resid {float(sc[0]):.2e} (oracle {T.solverResidual:.2e}) tdvp_err {float(sc[1]):.6e} (oracle {T.tdvp_error:.6e})", flush=True)
    big = np.abs(T.ev / T.ev[-1]) > 1e-8
    print(f"   snr rel err on well-conditioned modes {np.abs(snr[big] / T.snr[big] - 1).max():.1e}; |VtF| err {err(np.abs(VtF[big]), np.abs(T.VtF[big])):.1e}; rhoVar err {err(rhoVar[big], T.rhoVar[big]):.1e}", flush=True)

# cholesky
for n in (37, 300, 1000, 2053):
    ld = pad(n); A = rng.normal(size=(n + 50, n)); Sn = A.T @ A / n + 1e-3 * np.eye(n); Fn = rng.normal(size=n)
    S = torch.zeros(ld, ld, device=dev, dtype=torch.float64); S[:n, :n] = torch.tensor(Sn, device=dev); F = torch.tensor(Fn, device=dev)
    x = torch.zeros(n, device=dev, dtype=torch.float64); info = torch.zeros(1, device=dev, dtype=torch.int32)
    torch.cuda.synchronize(); t0 = time.time()
    _lib.check(L.vmcpde_chol_solve(_lib.ptr(S), n, ld, _lib.ptr(F), _lib.ptr(x), _lib.ptr(info), _lib.stream()))
    torch.cuda.synchronize()
    print(f"chol n={n}: err {err(x.cpu(), np.linalg.solve(Sn, Fn)):.1e} info {int(info)} sec {time.time() - t0:.4f}", flush=True)
