"""CPU restatement of vmc_fluids/exact_dyn.py (test infrastructure: checks vmcpde_particles_step; never imported by the
product).  Follows exact_dyn.py:31-47 (Hamiltonian field, uncoupled), :50-53 (fluid-dynamics-paper field), :56-62
(update_fun_phaseSpace), :65-67 (update_fun_Diff), :70-76 (integrate_single_coord), :79-82 (integrate: split(key, N),
per-particle split(., 4), normal(key, shape=coord.shape)).  Parity unpinned against JAX (see oracle/threefry.py)."""
import numpy as np

from . import threefry


def velocity_hamiltonian(x, p):
    v = np.zeros_like(x)
    v[0::2] = x[1::2] / p["m"]
    v[1::2] = -(p["m"] * p["omega"] ** 2 * x[0::2] + 4.0 * p["lam"] * x[0::2] ** 3)
    return v


def velocity_fluidpaper(x, p):
    c = np.cos(np.pi * p["t"] / p["T"])
    v = np.zeros_like(x)
    v[0] = -np.sin(np.pi * x[0]) ** 2 * np.sin(2 * np.pi * x[1]) * c
    v[1] = np.sin(np.pi * x[1]) ** 2 * np.sin(2 * np.pi * x[0]) * c
    return v


def update_phase_space(x, p, vel, dt, key):
    mask = np.zeros_like(x); mask[1::2] = 1.0
    v_adv = vel(x, p)
    v_diff = np.sqrt(2 * p["m"] * p["gamma"] * p["T"] / dt) * threefry.normal(key, x.shape[0])
    v_damp = -p["gamma"] * x
    return v_adv + v_diff * mask + v_damp * mask


def update_diffusion(x, p, vel, dt, key):
    return p["D"] * np.sqrt(2 / dt) * threefry.normal(key, x.shape[0])


def integrate_single(x, dt, p, vel, update, key):
    ks = threefry.split(key, 4)
    k1 = update(x, p, vel, dt / 6, ks[0])
    k2 = update(x + dt * 0.5 * k1, p, vel, dt / 3, ks[1])
    k3 = update(x + dt * 0.5 * k2, p, vel, dt / 3, ks[2])
    k4 = update(x + dt * k3, p, vel, dt / 6, ks[3])
    return x + dt * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0


def integrate(coords, dt, p, vel, update, key):
    keys = threefry.split(key, coords.shape[0])
    return np.stack([integrate_single(coords[i], dt, p, vel, update, keys[i]) for i in range(coords.shape[0])])
