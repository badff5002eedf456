"""Mirror of vmc_fluids/exact_dyn.py: the particle (Langevin / Hamiltonian) integrator the reference uses as an
independent check of the density dynamics, on the GPU.

`integrate(coords, dt, parameters, vel_field, update_fun, key)` keeps the reference's signature (exact_dyn.py:79-82):
`vel_field` and `update_fun` are the module-level functions below (they select the kernel's mode; arbitrary Python
callables cannot run inside the kernel and raise).  Noise is drawn exactly as the reference does: per-particle keys
`split(key, N)`, four stage keys `split(., 4)`, `normal(key, shape=coord.shape)` in JAX's threefry counter layout."""
import numpy as np
import torch

from . import _kernels, _capi, _threefry


def _velocity_field_hamiltonian(coord, evolParams):
    """exact_dyn.py:31-47, uncoupled branch: (dx/dt, dp/dt) = (p/m, -m w^2 x - 4 lam x^3) on interleaved (x, p)."""
    c = _kernels.as_dev(coord)
    v = torch.empty_like(c)
    v[..., 0::2] = c[..., 1::2] / evolParams["m"]
    v[..., 1::2] = -(evolParams["m"] * evolParams["omega"] ** 2 * c[..., 0::2] + 4.0 * evolParams["lam"] * c[..., 0::2] ** 3)
    return v


def _velocity_field_fluiddynpaper(coord, parameters):
    """exact_dyn.py:50-53."""
    c = _kernels.as_dev(coord)
    x, y = c[..., 0], c[..., 1]
    f = np.cos(np.pi * parameters["t"] / parameters["T"])
    return torch.stack([-torch.sin(np.pi * x) ** 2 * torch.sin(2 * np.pi * y) * f,
                        torch.sin(np.pi * y) ** 2 * torch.sin(2 * np.pi * x) * f], dim=-1)


def update_fun_phaseSpace(coord, parameters, vel_field, dt, key):
    """exact_dyn.py:56-62 (selector for `integrate`; the arithmetic runs in vmcpde_particles_step)."""
    raise NotImplementedError("update functions select the kernel mode of exact_dyn.integrate; call integrate(...)")


def update_fun_Diff(coord, parameters, vel_field, dt, key):
    """exact_dyn.py:65-67 (selector for `integrate`)."""
    raise NotImplementedError("update functions select the kernel mode of exact_dyn.integrate; call integrate(...)")


_UPDATES = {update_fun_phaseSpace: 0, update_fun_Diff: 1}
_FIELDS = {_velocity_field_hamiltonian: 0, _velocity_field_fluiddynpaper: 1, None: -1}


def integrate(coords, dt, parameters, vel_field, update_fun, key):
    """exact_dyn.py:79-82: one stochastic step of all particles; returns the new coordinates (N, d)."""
    if update_fun not in _UPDATES or vel_field not in _FIELDS:
        raise NotImplementedError("exact_dyn.integrate runs the reference's update functions / velocity fields only")
    eq = _capi.make_equation("diffusion", {k: parameters[k] for k in ("D", "m", "omega", "lam", "T", "gamma") if k in parameters},
                             t=parameters.get("t", 0.0))
    out = _kernels.as_dev(coords).clone()
    if out.ndim != 2:
        raise ValueError("coords must have shape (N, dim)")
    k = np.asarray(key, dtype=np.uint32).reshape(2) if not np.isscalar(key) else _threefry.PRNGKey(key)
    _kernels.particles_step(out, dt, _UPDATES[update_fun], _FIELDS[vel_field], eq, k)
    return out


def mc_integral(coords, lim=1):
    """exact_dyn.py:128-129: fraction of particles inside the ball of radius lim."""
    c = _kernels.as_dev(coords)
    return (torch.linalg.norm(c, dim=-1) < lim).sum() / c.shape[0]
