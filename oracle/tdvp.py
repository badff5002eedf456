"""Oracle restatement of evolutionEq.py, tdvp.py, stepper.py (+ mpi_wrapper reductions).

Follows: evolutionEq.py:18-45,53-119; tdvp.py:36-52,57-94,96-164; stepper.py:45-91,129-145;
mpi_wrapper.py:21-25,129-274 (single process: sums over the sample axis / globNumSamples).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import math
import numpy as np
import torch

from . import threefry

EQ_PARAMS = {  # evolutionEq.py:61-77
    "diffusion": {"D": 1.0},
    "diffusion_drift": {"D": 1.0, "mu": 4.0},
    "advection_hamiltonian": {"m": 1.0, "omega": 1.0, "lam": 0.0},
    "advection_hamiltonian_wDiss": {"m": 1.0, "omega": 1.0, "T": 10.0, "gamma": 1.0, "lam": 0.0},
    "advection_paper": {"T": 5.0},
}


def random_D_factor(dim):
    """evolutionEq.py:18-20: A = normal(PRNGKey(0), (dim, dim)); D = A.T @ A."""
    return threefry.normal(threefry.prng_key(0), dim * dim).reshape(dim, dim)


def velocity_hamiltonian(x, p):
    """evolutionEq.py:30-45 (uncoupled branch): v = J grad H on interleaved (x,p) coordinates."""
    xs, ps = x[:, 0::2], x[:, 1::2]
    v = torch.zeros_like(x)
    v[:, 0::2] = ps / p["m"]
    v[:, 1::2] = -(p["m"] * p["omega"] ** 2 * xs + 4.0 * p["lam"] * xs ** 3)
    return v


def velocity_paper(x, t, p):
    """evolutionEq.py:23-27."""
    X, Y = x[:, 0], x[:, 1]
    c = math.cos(math.pi * t / p["T"])
    return torch.stack([-torch.sin(math.pi * X) ** 2 * torch.sin(2 * math.pi * Y) * c,
                        torch.sin(math.pi * Y) ** 2 * torch.sin(2 * math.pi * X) * c], dim=1)


def local_terms(state, x, name, t=0.0):
    """EvolutionEquation.__call__ (evolutionEq.py:81-119) -> (Eloc, O, logp, grad_x)."""
    logp, g, O = state.eval_coordgrads(x)
    x = torch.as_tensor(x)
    if name == "diffusion":
        H = state.hessian(x)
        E = EQ_PARAMS[name]["D"] * ((g * g).sum(-1) + torch.einsum("aii->a", H))
    elif name == "diffusion_drift":
        H = state.hessian(x)
        E = EQ_PARAMS[name]["D"] * ((g * g).sum(-1) + torch.einsum("aii->a", H)) + EQ_PARAMS[name]["mu"] * g.sum(-1)
    elif name == "diffusion_anisotropic":
        A = torch.as_tensor(random_D_factor(x.shape[1]))
        D = A.T @ A
        H = state.hessian(x)
        E = torch.einsum("ai,ij,aj->a", g, D, g) + torch.einsum("aij,ji->a", H, D)
    elif name == "advection_hamiltonian":
        E = -(g * velocity_hamiltonian(x, EQ_PARAMS[name])).sum(-1)
    elif name == "advection_paper":
        E = -(g * velocity_paper(x, t, EQ_PARAMS[name])).sum(-1)
    elif name == "advection_hamiltonian_wDiss":
        p = EQ_PARAMS[name]
        H = state.hessian(x)
        adv = -(g * velocity_hamiltonian(x, p)).sum(-1)
        diff = p["m"] * p["gamma"] * p["T"] * ((g[:, 1::2] ** 2).sum(-1) + torch.einsum("aii->a", H[:, 1::2, 1::2]))
        damp = p["gamma"] * (x[:, 1::2] * g[:, 1::2]).sum(-1)
        E = adv + diff + damp
    else:
        raise KeyError(name)
    return E, O, logp, g


class OracleTDVP:
    """tdvp.py:20-94 on host float64 (numpy)."""

    def __init__(self, useSNR=False, snrTol=2.0, svdTol=1e-11, diagonalShift=0.0):
        self.useSNR, self.snrTol, self.svdTol, self.diagonalShift = useSNR, snrTol, svdTol, diagonalShift

    def equation(self, Eloc, O, logp, n_glob=None):
        """get_tdvp_equation, tdvp.py:36-52."""
        E = np.asarray(Eloc, dtype=np.float64)
        O = np.asarray(O, dtype=np.float64)
        lp = np.asarray(logp, dtype=np.float64)
        N = n_glob or E.shape[0]
        self.N = N
        self.ElocMean = E.sum() / N
        self.ElocMeanAbs = np.abs(E).sum() / N
        self.ElocVar = ((E - self.ElocMean) ** 2).sum() / N
        dE = E - self.ElocMean
        self.gradMean = O.sum(0) / N
        dO = O - self.gradMean
        EO = dE[:, None] * dO
        self.F0 = EO.sum(0) / N
        self.S0 = dO.T @ dO / N
        w = lp[:, None] * dO
        self.SExp = w.T @ w / N
        S = self.S0
        if self.diagonalShift > 1e-10:
            S = S + np.diag(self.diagonalShift * np.diag(S))
        return S, self.F0, EO

    def solve(self, Eloc, O, logp, n_glob=None):
        """solve + transform_to_eigenbasis, tdvp.py:57-94."""
        self.S, F, EO = self.equation(Eloc, O, logp, n_glob)
        N = self.N
        self.ev, self.V = np.linalg.eigh(self.S)
        self.VtF = self.V.T @ F
        EOv = EO @ self.V
        m = EOv.sum(0) / N
        self.rhoVar = ((EOv - m) ** 2).sum(0) / N
        with np.errstate(divide="ignore", invalid="ignore"):
            self.snr = np.sqrt(np.abs(N * self.VtF * self.VtF / self.rhoVar))
            r = np.abs(self.ev / self.ev[-1])
            self.invEv = np.where(r > 1e-14, 1.0 / self.ev, 0.0)
            reg = 1.0 / (1.0 + (self.svdTol / r) ** 6)
            if self.useSNR:
                reg = reg * 1.0 / (1.0 + (self.snrTol / self.snr) ** 6)
        self.regularizer = reg
        update = self.V @ (self.invEv * reg * self.VtF)
        E = np.asarray(Eloc, dtype=np.float64)
        self.tdvp_error = 1.0 + (update @ self.S0 @ update - 2.0 * self.F0 @ update) / np.mean(E ** 2)
        self.solverResidual = np.linalg.norm(self.S @ update - F) / np.linalg.norm(F)
        return update

    def rhs(self, state, theta, name, n_samples, t=0.0, observables=True):
        """TDVP.__call__ (tdvp.py:96-164) for nSamplesObs <= nSamplesTDVP."""
        state.theta = torch.as_tensor(np.asarray(theta, dtype=np.float64)).clone()
        x, lp_s, _ = state.sample(n_samples)
        E, O, lp, _ = local_terms(state, x, name, t)
        upd = self.solve(E.numpy(), O.numpy(), lp.numpy())
        info = observables_info(state, x.numpy(), lp.numpy(), E.numpy()) if observables else {}
        return upd, info


def observables_info(state, x, logp, Eloc, n_obs=None):
    """tdvp.py:143-162."""
    info = {}
    mean = x.mean(0)
    info["x1"] = mean
    info["covar"] = np.cov(x.T, ddof=0)
    info["entropy"] = -logp.mean()
    for m in (3, 4, 5, 6):
        info[f"x{m}"] = ((x - mean) ** m).mean(0)
    info["max_grad"] = Eloc.max()
    n = n_obs or x.shape[0]
    d = x.shape[1]
    # tdvp.py:154-155: both draws use psi.sampler.key itself (not a split of it)
    s = threefry.normal(state.key, n * d).reshape(n, d)
    u = threefry.uniform(state.key, n)
    s = s / np.linalg.norm(s, axis=-1, keepdims=True) * u[:, None] ** (1.0 / d)
    from scipy.special import gamma
    for lim in (1, 0.5, 0.1):
        lim_s = lim * math.sqrt(10.0)
        vol = math.pi ** (d / 2) / gamma(d / 2 + 1) * lim_s ** d
        info[f"integral_{lim}sigma"] = float(np.exp(state.logp(lim_s * s).numpy()).mean() * vol)
    return info


# ------------------------------------------------------------------------ steppers (stepper.py)
def heun_step(f, y, dt, max_step, increase_fac):
    """FixedStepper.step mode 'Heun', stepper.py:129-139.  f(y, k) -> update."""
    dt = min(dt * increase_fac, max_step)
    k0 = f(y, 0)
    k1 = f(y + dt * k0, 1)
    return y + 0.5 * dt * (k0 + k1), dt


def euler_step(f, y, dt, max_step, increase_fac):
    """stepper.py:141-145."""
    dt = min(dt * increase_fac, max_step)
    return y + dt * f(y, 0), dt


def adaptive_heun_step(f, y, dt, tol, max_step, norm):
    """AdaptiveHeun.step, stepper.py:45-91.  norm(v) = normFunction(v, f.SExp) evaluated after the 5th call."""
    fe = 0.5
    while fe < 1.0:
        k0 = f(y, 0)
        k1 = f(y + dt * k0, 1)
        dy0 = 0.5 * dt * (k0 + k1)
        k10 = f(y + 0.5 * dt * k0, 2)
        dy1 = 0.25 * dt * (k0 + k10)
        k01 = f(y + dy1, 3)
        k11 = f(y + dy1 + 0.5 * dt * k01, 4)
        dy1 = dy1 + 0.25 * dt * (k01 + k11)
        fe = tol / norm(dy1 - dy0)
        fac = min(max(0.9 * fe ** 0.33333, 0.2), 2.0)
        real_dt = dt
        dt = min(dt * fac, max_step)
    return y + dy1, real_dt, dt
