"""NumPy prototype of the blocked (panel) tridiagonalisation + compact-WY back-transform implemented in
csrc/eigh_blocked.cu.  Row convention of eigh.cu: reflector v_j lives in row j, columns j+1.. (v_j[j+1] = 1).
Used to validate the algebra (redundant w(j+1), p.v from partial dots, masked trailing update) before the
CUDA version; run by hand."""
import numpy as np


def tridiag_blocked(A, nb=8):
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n); e = np.zeros(n); tau = np.zeros(n)
    Vst = np.zeros((n, n))          # reflector store: Vst[j, c] = v_j(c)
    Gst = np.zeros((n, n))          # Gst[j, k] = v_k . v_j for k < j in the same panel (by-product)
    for j0 in range(0, n, nb):
        nbp = min(nb, n - j0)
        Vp = np.zeros((nbp, n)); Wp = np.zeros((nbp, n))
        # prologue: acol = row j0
        acol = A[j0].copy()
        for i in range(nbp):
            j = j0 + i
            d[j] = acol[j]
            m = n - j - 1
            if m >= 2:
                alpha = acol[j + 1]
                sigma = np.sum(acol[j + 2:] ** 2)
                if sigma == 0.0:
                    t = 0.0; beta = alpha
                    v = np.zeros(n); v[j + 1] = 1.0
                else:
                    beta = -np.copysign(np.sqrt(alpha * alpha + sigma), alpha)
                    t = (beta - alpha) / beta
                    v = np.zeros(n); v[j + 2:] = acol[j + 2:] / (alpha - beta); v[j + 1] = 1.0
                e[j] = beta; tau[j] = t
                Vst[j] = v
                Vp[i] = v
                # phase A: y = A22 v (un-updated trailing matrix), partial dots
                y = np.zeros(n)
                y[j + 1:] = A[j + 1:, j + 1:] @ v[j + 1:]
                PV = Vp[:i] @ v; PW = Wp[:i] @ v; YV = y @ v
                Gst[j, j0:j0 + i] = PV
                # phase B
                pv = t * (YV - 2.0 * PV @ PW)
                yc = y - Vp[:i].T @ PW - Wp[:i].T @ PV
                w = t * yc - 0.5 * t * pv * v
                w[:j + 1] = 0.0
                # check redundant formula for w(j+1)
                w1 = t * (y[j + 1] - Vp[:i, j + 1] @ PW - Wp[:i, j + 1] @ PV) - 0.5 * t * pv
                assert abs(w1 - w[j + 1]) <= 1e-12 * (1 + abs(w1))
                Wp[i] = w
            elif m == 1:
                e[j] = acol[j + 1]; tau[j] = 0.0
            # next column prep (if inside the panel)
            if i + 1 < nbp:
                r = j + 1
                acol = A[r].copy()
                acol -= Vp[:i + 1, r] @ Wp[:i + 1] + Wp[:i + 1, r] @ Vp[:i + 1]
        j1 = j0 + nbp
        if j1 < n:
            X = np.concatenate([Vp, Wp]); Y = np.concatenate([Wp, Vp])
            X[:, :j1] = 0; Y[:, :j1] = 0
            ja = j1 // 4 * 4  # aligned-down region start (emulates the 128 alignment)
            A[ja:, ja:] -= X[:, ja:].T @ Y[:, ja:]
    return d, e, tau, Vst, Gst


def backtransform_blocked(Z, Vst, tau, kb=8, G=None):
    """V = H_0 ... H_{n-3} Z with blocks of kb reflectors: Q_B = I - Y T Y^T."""
    n = Z.shape[0]
    Z = Z.copy()
    nblk = (n + kb - 1) // kb
    for b in range(nblk - 1, -1, -1):
        ja, jb = b * kb, min(n, (b + 1) * kb)
        Y = Vst[ja:jb].T                       # n x kb
        Gm = Y.T @ Y
        k = jb - ja
        T = np.zeros((k, k))
        for i in range(k):
            T[i, i] = tau[ja + i]
            if i > 0:
                T[:i, i] = -tau[ja + i] * (T[:i, :i] @ Gm[:i, i])
        G1 = Y.T @ Z
        G2 = T @ G1
        Z -= Y @ G2
    return Z


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 5, 17, 64, 101):
        B = rng.normal(size=(n, n)); A = (B + B.T) / 2
        if n == 17:
            A[3, 5:] = 0; A[5:, 3] = 0  # a sigma == 0 column
        d, e, tau, Vst, Gst = tridiag_blocked(A, nb=8)
        T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
        lam, Z = np.linalg.eigh(T)
        ref = np.linalg.eigvalsh(A)
        V = backtransform_blocked(Z, Vst, tau, kb=8)
        print(n, np.abs(lam - ref).max(), np.abs(A @ V - V * lam).max(), np.abs(V.T @ V - np.eye(n)).max())
