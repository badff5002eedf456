"""Sizing study (CPU, NumPy; not part of the product): how far would a rank-revealing compression of S move the TDVP update?

DESIGN.md section 7: the serial eigen-solve is what caps multi-GPU scaling, and S = <dO dO^T> has a structural null space
(58-63 % of the eigenvalues of the reference's stored runs sit below 1e-8 lambda_max).  Route studied here, against the
reference's regularised solve (tdvp.py:57-94, restated in oracle/tdvp.py):

    pivoted Cholesky  S ~= L L^T  stopped at relative pivot `tol`  (rank r << P)
    eigen-solve of the r x r matrix  L^T L = W diag(mu) W^T          (same non-zero spectrum as L L^T)
    V_r = L W diag(mu)^(-1/2),   update = V_r diag(reg(mu) / mu) V_r^T F

usage: python tools/proto/rank_revealing_solve.py            (d = 6, P = 411, different_add: the stored INN run's architecture)
"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import flow, tdvp  # noqa: E402


def pivoted_cholesky(S, tol):
    """Diagonal-pivoted Cholesky; stops when the largest remaining pivot falls below tol * the first one."""
    n = S.shape[0]
    d = np.diag(S).copy()
    perm = np.arange(n)
    L = np.zeros((n, n))
    d0 = d.max()
    r = 0
    for k in range(n):
        j = k + int(np.argmax(d[perm[k:]]))
        if d[perm[j]] <= tol * d0:
            break
        perm[[k, j]] = perm[[j, k]]
        p = perm[k]
        L[p, k] = np.sqrt(d[p])
        rest = perm[k + 1:]
        L[rest, k] = (S[rest, p] - L[rest, :k] @ L[p, :k]) / L[p, k]
        d[rest] -= L[rest, k] ** 2
        r = k + 1
    return L[:, :r], r


def main():
    d, depth, h, N = 6, 4, 3, 10000
    ups, downs, key = flow.make_index_splits(d, depth, 1)
    spec = flow.FlowSpec(dim=d, depth=depth, hidden=(h,), variant="different_add", offset=np.array([1., 0, 0, 1, 0, 0]),
                         inds_up=ups, inds_down=downs)
    th = flow.init_params_flax(spec, key)
    st = flow.OracleState(spec, th)
    x, _, _ = st.sample(N)
    E, O, lp, _ = tdvp.local_terms(st, x, "advection_hamiltonian_wDiss", 0.0)
    T = tdvp.OracleTDVP()
    upd = T.solve(E.numpy(), O.numpy(), lp.numpy())
    S, F = T.S, T.F0
    ev = T.ev
    P = S.shape[0]
    print(f"P = {P}, lambda_max = {ev[-1]:.3e}; eigenvalues above 1e-8 / 1e-11 / 1e-14 of lambda_max: "
          f"{int((ev > 1e-8 * ev[-1]).sum())} / {int((ev > 1e-11 * ev[-1]).sum())} / {int((ev > 1e-14 * ev[-1]).sum())}")
    print(f"reference solve: tdvp_error = {T.tdvp_error:.6e}, residual = {T.solverResidual:.2e}, |update|_S = {np.sqrt(upd @ T.S0 @ upd):.6e}")
    for tol in (1e-10, 1e-12, 1e-14, 1e-16):
        L, r = pivoted_cholesky(S, tol)
        mu, W = np.linalg.eigh(L.T @ L)
        keep = mu > 0
        mu, W = mu[keep], W[:, keep]
        Vr = L @ W / np.sqrt(mu)
        rel = np.abs(mu / mu[-1])
        reg = 1.0 / (1.0 + (T.svdTol / rel) ** 6)
        u2 = Vr @ (reg / mu * (Vr.T @ F))
        du = u2 - upd
        err_S = np.sqrt(du @ T.S0 @ du) / np.sqrt(upd @ T.S0 @ upd)
        te = 1.0 + (u2 @ T.S0 @ u2 - 2.0 * F @ u2) / np.mean(E.numpy() ** 2)
        top = ev[::-1][:len(mu)]
        big = top > 1e-8 * ev[-1]
        ev_err = np.abs(mu[::-1][:len(top)][big] / top[big] - 1).max()
        orth = np.abs(Vr.T @ Vr - np.eye(Vr.shape[1])).max()
        print(f"tol {tol:.0e}: rank {r:4d} ({100 * r / P:4.1f} % of P)  |d update|_S / |update|_S = {err_S:.2e}  "
              f"tdvp_error {te:.6e} (diff {te - T.tdvp_error:+.1e})  max rel. error of ev above 1e-8 lambda_max {ev_err:.1e}  |Vr^T Vr - I| {orth:.1e}")


if __name__ == "__main__":
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    main()
