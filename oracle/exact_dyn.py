"""CPU restatement of vmc_fluids/exact_dyn.py (test infrastructure: checks vmcpde_particles_step; never imported by the
product).  Follows exact_dyn.py:31-47 (Hamiltonian field, uncoupled), :50-53 (fluid-dynamics-paper field), :56-62
(update_fun_phaseSpace), :65-67 (update_fun_Diff), :70-76 (integrate_single_coord), :79-82 (integrate: split(key, N),
per-particle split(., 4), normal(key, shape=coord.shape)).

Parity PINNED on reference-held data: `integrate_batch` (the vectorised statement of `integrate`) reproduces every record
of the reference's stored run paper_plot/data_phaseSpace/Wiener/Nsamples10000_T10.0/infos.hdf5 (its own
exact_dyn.py:85-153 main loop, N=10^4, dt=1e-2) to round-off (tests/test_reference_pins.py, tests/golden/ref_wiener_T10.npz)."""
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


def integrate_batch(coords, dt, p, key, mode="phase_space"):
    """Vectorised `integrate` for the Hamiltonian field (the per-particle loop above, all particles at once)."""
    coords = np.asarray(coords, dtype=np.float64)
    n, d = coords.shape
    ks = threefry.split_each(threefry.split(key, n), 4)
    mask = np.zeros(d); mask[1::2] = 1.0

    def vel(x):
        v = np.zeros_like(x)
        v[:, 0::2] = x[:, 1::2] / p["m"]
        v[:, 1::2] = -(p["m"] * p["omega"] ** 2 * x[:, 0::2] + 4.0 * p["lam"] * x[:, 0::2] ** 3)
        return v

    def update(x, h, keys):
        if mode == "diffusion":
            return p["D"] * np.sqrt(2 / h) * threefry.normal_each(keys, d)
        return vel(x) + np.sqrt(2 * p["m"] * p["gamma"] * p["T"] / h) * threefry.normal_each(keys, d) * mask - p["gamma"] * x * mask

    k1 = update(coords, dt / 6, ks[:, 0])
    k2 = update(coords + dt * 0.5 * k1, dt / 3, ks[:, 1])
    k3 = update(coords + dt * 0.5 * k2, dt / 3, ks[:, 2])
    k4 = update(coords + dt * k3, dt / 6, ks[:, 3])
    return coords + dt * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0


def reference_main_loop(n_steps, n=10000, dim=6, offset=(1, 0, 1, 0, 1, 0), dt=1e-2, T=10.0, step=None):
    """exact_dyn.py:85-153 ("hamiltonian" case): records (x1, covar, three ball fractions) BEFORE each step.
    `step(coords, dt, p, key)` replaces the integrator (the GPU test passes the device one)."""
    p = {"T": T, "t": 0.0, "gamma": 1.0, "m": 1.0, "omega": 1.0, "lam": 0.0}
    coords = threefry.normal(threefry.prng_key(0), n * dim).reshape(n, dim) + np.asarray(offset, dtype=np.float64)
    key = threefry.prng_key(0)
    rec = {"x1": [], "covar": [], "integral_1sigma": [], "integral_0.5sigma": [], "integral_0.1sigma": []}
    step = step or (lambda c, h, pp, k: integrate_batch(c, h, pp, k))
    for _ in range(n_steps):
        ks = threefry.split(key)
        key, use = ks[0], ks[1]
        c = np.asarray(coords)
        rec["x1"].append(c.mean(0))
        rec["covar"].append(np.cov(c.T, ddof=0))
        r = np.linalg.norm(c, axis=-1)
        for lim in (1, 0.5, 0.1):
            rec[f"integral_{lim}sigma"].append(np.sum(r < lim * np.sqrt(T)) / n)
        coords = step(coords, dt, p, use)
    return {k: np.array(v) for k, v in rec.items()}
