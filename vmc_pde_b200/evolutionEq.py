"""Mirror of vmc_fluids/evolutionEq.py: the PDE operators giving the local time derivative of log p.

`EvolutionEquation.__call__(vState, configs, t)` returns (time_grads (1,N), param_grads (1,N,P), logProbs (1,N))
like the reference (evolutionEq.py:81-119), computed by ONE fused kernel (forward jets + reverse sweep) instead of
value_and_grad + jacrev(jacfwd) + einsum.  TDVP.__call__ uses `equation_struct()` to run the same kernel chunk by
chunk without materialising param_grads for all samples.
"""
from dataclasses import dataclass
import numpy as np
import torch

from . import _kernels, _capi, _threefry, global_defs


def _getRandomD_factor(dim):
    """evolutionEq.py:18-20: A = normal(PRNGKey(0), (dim, dim)); D = A.T @ A.  Returns A (device, dim x dim)."""
    return _kernels.normal(_threefry.PRNGKey(0), 0, dim * dim, dim * dim).view(dim, dim).contiguous()


def _getRandomD_matrix(dim):
    A = _getRandomD_factor(dim)
    return A.T @ A


def _velocity_field_MLPaper(evolParams, coord, t):
    """evolutionEq.py:23-27 (host helper kept for plotting parity; the kernels evaluate the field on the device)."""
    x, y = coord[0], coord[1]
    c = np.cos(np.pi * t / evolParams["T"])
    return np.array([-np.sin(np.pi * x) ** 2 * np.sin(2 * np.pi * y) * c, np.sin(np.pi * y) ** 2 * np.sin(2 * np.pi * x) * c])


def _velocity_field_hamiltonian(evolParams, coord, t):
    """evolutionEq.py:30-45, uncoupled branch: (dx/dt, dp/dt) = (p/m, -m w^2 x - 4 lam x^3) on interleaved coords."""
    coord = np.asarray(coord, dtype=np.float64)
    v = np.zeros_like(coord)
    v[0::2] = coord[1::2] / evolParams["m"]
    v[1::2] = -(evolParams["m"] * evolParams["omega"] ** 2 * coord[0::2] + 4 * evolParams["lam"] * coord[0::2] ** 3)
    return v


@dataclass
class EvolutionEquation:
    """evolutionEq.py:48-119."""
    dim: int = 2
    name: str = "diffusion"

    def __post_init__(self, eqParams={"D": 1.}):
        self.function_dict = {"diffusion": self._diffusion_eq,
                              "diffusion_drift": self._diffusion_eq_wDrift,
                              "diffusion_anisotropic": self._diffusion_eq_anisotropic,
                              "advection_paper": self._advection,
                              "advection_hamiltonian": self._advection,
                              "advection_hamiltonian_wDiss": self._advection_wDiss,
                              }
        if self.name not in self.function_dict:
            raise KeyError(self.name)
        self._A = _getRandomD_factor(self.dim) if self.name == "diffusion_anisotropic" else None
        # evolutionEq.py:61-77
        self.eqParams = {"diffusion": {"D": 1},
                         "diffusion_anisotropic": {"D": (self._A.T @ self._A) if self._A is not None else None},
                         "diffusion_drift": {"D": 1, "mu": 4},
                         "advection_paper": {"params": {"T": 5}, "vel_field": _velocity_field_MLPaper},
                         "advection_hamiltonian": {"params": {"m": 1.0, "omega": 1.0, "lam": 0.0},
                                                   "vel_field": _velocity_field_hamiltonian},
                         "advection_hamiltonian_wDiss": {"params": {"m": 1.0, "omega": 1.0, "T": 10.0, "gamma": 1.0, "lam": 0.0},
                                                         "vel_field": _velocity_field_hamiltonian}}

    def equation_struct(self, t=0.0):
        """The vmcpde_equation for this operator at time t."""
        p = self.eqParams[self.name]
        flat = dict(p.get("params", {}))
        for k in ("D", "mu"):
            if k in p and not isinstance(p[k], torch.Tensor) and p[k] is not None:
                flat[k] = p[k]
        tang = self._A.data_ptr() if self._A is not None else None
        return _capi.make_equation(self.name, flat, t, tang)

    def __call__(self, *args):
        return self.function_dict[self.name](*args)

    def _fused(self, vState, configs, t):
        c = vState._coords(configs)
        nd, nb = c.shape[0], c.shape[1]
        h = vState.net.handle
        n = nd * nb
        O = _kernels.empty(max(n, 1), h.Pp)
        out = _kernels.local_terms(h, vState._flat, c.reshape(n, self.dim), self.equation_struct(t), O=O, ldo=h.Pp,
                                   want=("eloc", "logp"))
        return out["eloc"].view(nd, nb), O[:n, :h.P].view(nd, nb, h.P), out["logp"].view(nd, nb)

    # the six reference entry points share the fused kernel; the mode is carried by equation_struct()
    def _diffusion_eq(self, vState, configs, t):              # evolutionEq.py:84-87
        return self._fused(vState, configs, t)

    def _diffusion_eq_wDrift(self, vState, configs, t):       # evolutionEq.py:89-94
        return self._fused(vState, configs, t)

    def _diffusion_eq_anisotropic(self, vState, configs, t):  # evolutionEq.py:96-100
        return self._fused(vState, configs, t)

    def _advection(self, vState, configs, t):                 # evolutionEq.py:102-105
        return self._fused(vState, configs, t)

    def _advection_wDiss(self, vState, configs, t):           # evolutionEq.py:107-119
        return self._fused(vState, configs, t)
