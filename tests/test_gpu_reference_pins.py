"""The CUDA path against REFERENCE-HELD data (tests/golden/ref_*.npz = stored outputs of the reference's own JAX runs,
see tests/golden/make_reference_pins.py and the CPU twin tests/test_reference_pins.py).

  * device RNG + particle kernel reproduce ALL 1201 stored records of the reference's exact_dyn run to round-off;
  * VarState.init_net reproduces flax's initialisation stream (equal to the restated one, which reproduces the stored
    spectrum); the first records of the two stored Gauss-latent TDVP runs are reproduced through the main.py call sequence;
  * the TDVP-evolved density of the checked-in 'harmonicOsc_diff' physics follows the stored particle trajectory and ends
    at the constants the reference plots (paper_plot_phaseSpaceTempDifference.py:87,129-131).
"""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flow as oflow, threefry as othreefry, exact_dyn as oexact

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def norm_fun(v, S):  # main.py:24-26
    return v @ S @ v


def build(d, depth, h, variant, eqname, offset):
    from vmc_pde_b200 import sampler, var_state, evolutionEq, net
    smp = sampler.Sampler(dim=d, numChains=30, name="Gauss", mcmc_info={"offset": offset, "bound": 0.25})
    net.SingleBlock.different_add = (variant == "different_add")
    try:
        vs = var_state.VarState(smp, d, 1, depth, network_args={"intmediate": (h,), "offset": offset, "latentSpaceName": "Gauss", "dim": d})
    finally:
        net.SingleBlock.different_add = False
    return smp, vs, evolutionEq.EvolutionEquation(dim=d, name=eqname)


def test_device_rng_and_particle_kernel_reproduce_every_stored_record():
    """exact_dyn.py:85-153 with the device integrator: 1201 records of x1, covar (1e-10) and the ball counts (exact)."""
    from vmc_pde_b200 import exact_dyn, _kernels
    g = load("ref_wiener_T10")
    n, dim = 10000, 6
    key0 = np.array([0, 0], dtype=np.uint32)
    z = _kernels.normal(key0, 0, n * dim, n * dim).view(n, dim)                      # normal(PRNGKey(0), (N, 6))
    coords = z + _kernels.as_dev(np.array([1.0, 0, 1, 0, 1, 0]))
    p = {"T": 10.0, "t": 0.0, "gamma": 1.0, "m": 1.0, "omega": 1.0, "lam": 0.0}
    key = othreefry.prng_key(0)
    nrec = len(g["times"])
    x1 = torch.empty(nrec, dim, dtype=torch.float64, device=coords.device)
    cov = torch.empty(nrec, dim, dim, dtype=torch.float64, device=coords.device)
    balls = torch.empty(nrec, 3, dtype=torch.float64, device=coords.device)
    for i in range(nrec):
        ks = othreefry.split(key)
        key, use = ks[0], ks[1]
        m = coords.mean(0)
        dc = coords - m
        x1[i], cov[i] = m, dc.T @ dc / n
        r = torch.linalg.norm(coords, dim=-1)
        for j, lim in enumerate((1, 0.5, 0.1)):
            balls[i, j] = (r < lim * np.sqrt(10.0)).sum() / n
        coords = exact_dyn.integrate(coords, 1e-2, p, exact_dyn._velocity_field_hamiltonian, exact_dyn.update_fun_phaseSpace, use)
    assert np.abs(x1.cpu().numpy() - g["x1"]).max() < 1e-10
    assert np.abs(cov.cpu().numpy() - g["covar"]).max() < 1e-10
    for j, lim in enumerate((1, 0.5, 0.1)):      # counts / N: equal except where a particle sits within round-off of a sphere
        diff = np.abs(balls[:, j].cpu().numpy() - g[f"integral_{lim}sigma"])
        assert diff.max() < 1.01e-4 and int((diff > 1e-12).sum()) <= 3, (lim, diff.max(), int((diff > 1e-12).sum()))


def test_init_net_reproduces_the_flax_stream_and_first_stored_records():
    """VarState.init_net (var_state.py:110-124) == the restated flax stream; then main.py's call sequence against the
    first stored record of the d=8 diffusion run (first right-hand side) and of the d=6 phase-space run (second
    right-hand side of the first Heun step, incl. the 50 largest eigenvalues of S)."""
    from vmc_pde_b200 import tdvp, stepper
    # ---- d = 8, 'diffusion', Gauss latent
    g = load("ref_diff8_gauss")
    smp, vs, eq = build(8, 4, 4, "no_add", "diffusion", np.zeros(8))
    ups, downs, key = oflow.make_index_splits(8, 4, 1)
    spec = oflow.FlowSpec(dim=8, depth=4, hidden=(4,), variant="no_add", inds_up=ups, inds_down=downs)
    assert vs.net.inds_up == ups and vs.net.inds_down == downs
    assert np.array_equal(vs.get_parameters().cpu().numpy(), oflow.init_params_flax(spec, key))
    T = tdvp.TDVP()
    upd, info = T(vs.get_parameters(), 0.0, psi=vs, evolutionEq=eq, nSamplesTDVP=10000, nSamplesObs=10000, timings=None)
    f = lambda v: np.asarray(v.cpu())
    assert np.abs(f(info["x1"]) - g["x1"][0]).max() < 5e-5 and np.abs(f(info["covar"]) - g["covar"][0]).max() < 2e-4
    assert abs(float(info["entropy"]) - g["entropy"][0]) < 1e-4 and abs(float(info["max_grad"]) / g["max_grad"][0] - 1) < 5e-3
    for lim in (1, 0.5, 0.1):
        assert abs(float(info[f"integral_{lim}sigma"]) / g[f"integral_{lim}sigma"][0] - 1) < 2e-4
    # the stored run has 28 more parameters than the checked-in architecture (P = 392 vs 364): spectra only in bands
    assert abs(float(T.ev[-1]) / g["ev"][0][-1] - 1) < 0.03 and float(T.solverResidual) < 1e-9
    assert 0.3 < float(T.tdvp_error) / g["tdvp_error"][0] < 3.0
    # ---- d = 6, 'harmonicOsc_diff' (main.py:35,108-118), P = 411
    g = load("ref_inn_Tdiff")
    off = np.array([1.0, 0, 0, 1, 0, 0])
    smp, vs, eq = build(6, 4, 3, "different_add", "advection_hamiltonian_wDiss", off)
    assert vs.numParameters == 411
    st = stepper.FixedStepper(timeStep=1e-4, mode='Heun', maxStep=1e-2, increase_fac=1.3)
    T = tdvp.TDVP()
    dp, dt, info = st.step(0, T, vs.get_parameters(), evolutionEq=eq, psi=vs, nSamplesTDVP=10000, nSamplesObs=10000,
                           normFunction=norm_fun, timings=None, integrals=False)
    assert abs(dt - 1.3e-4) < 1e-18
    assert np.abs(f(info["x1"]) - g["x1"][0]).max() < 1e-3 and np.abs(f(info["covar"]) - g["covar"][0]).max() < 5e-3
    assert abs(float(info["entropy"]) - g["entropy"][0]) < 5e-3
    for lim in (1, 0.5, 0.1):
        assert abs(float(info[f"integral_{lim}sigma"]) / g[f"integral_{lim}sigma"][0] - 1) < 2e-3
    ev = f(T.ev)
    assert np.abs(ev[-50:] / g["ev"][0][-50:] - 1).max() < 1e-2
    cut = lambda e: int(np.sum(np.abs(e / e[-1]) < 1e-11))
    assert abs(cut(ev) - cut(g["ev"][0])) <= 25
    # snr of the dominant modes has the stored magnitude (it involves the force, i.e. the run's edited physics: band only)
    ratio = f(T.snr)[-6:] / g["snr"][0][-6:]
    assert 0.25 < ratio.min() and ratio.max() < 4.0, ratio


def test_tdvp_evolution_follows_the_stored_particle_run_and_ends_at_the_plotted_constants():
    """main.py's loop (Heun, dt = 1e-4 * 1.3^k <= 1e-2, N = 10^4, P = 411, 1214 steps to t = 12) on the checked-in
    damped-oscillator physics with the initial state of the stored particle run (offset [1,0,1,0,1,0]).  The evolved density
    must follow the reference's own particle trajectory (its N = 10^4 particles scatter by ~1.5 % in the variances) and
    reach the steady state the reference plots: entropy 6 * 1/2 log(2 pi e 10) and the three ball integrals
    0.0143877 / 2.96478e-4 / 2.07554e-8 (paper_plot_phaseSpaceTempDifference.py:87,129-131).
    Measured on a B200: means within 0.10, variances within 5 %, entropy within 0.04, end integrals within 0.3 %."""
    from vmc_pde_b200 import tdvp, stepper
    g = load("ref_wiener_T10")
    off = np.array([1.0, 0, 1, 0, 1, 0])
    smp, vs, eq = build(6, 4, 3, "different_add", "advection_hamiltonian_wDiss", off)
    st = stepper.FixedStepper(timeStep=1e-4, mode='Heun', maxStep=1e-2, increase_fac=1.3)
    T = tdvp.TDVP()
    t, checks, hist = 0.0, [0.25, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, 8.0, 12.0], {}
    evs = []
    while t < 12.0 + 1e-9:
        dp, dt, info = st.step(0, T, vs.get_parameters(), evolutionEq=eq, psi=vs, nSamplesTDVP=10000, nSamplesObs=10000,
                               normFunction=norm_fun, timings=None, integrals=False)
        vs.set_parameters(dp)
        evs.append(T.ev)
        # the logged record belongs to the trial state at t + dt (second right-hand side, stepper.py:136)
        if checks and t + dt >= checks[0]:
            hist[checks.pop(0)] = (t + dt, {k: v.cpu().numpy() for k, v in info.items()})
        t += dt
    assert not torch.equal(evs[0], evs[-1]) and float(T.solverResidual) < 1e-9     # histories do not alias (main.py:187)
    tw = g["times"]
    for tc, (tt, info) in hist.items():
        i = int(np.argmin(np.abs(tw - tt)))
        Cw = g["covar"][i]
        assert np.abs(info["x1"] - g["x1"][i]).max() < 0.15, (tc, info["x1"], g["x1"][i])
        assert np.abs(np.diag(info["covar"]) / np.diag(Cw) - 1).max() < 0.08, (tc, np.diag(info["covar"]), np.diag(Cw))
        assert abs(info["covar"][0, 1] - Cw[0, 1]) < 0.45
        ent_w = 0.5 * np.linalg.slogdet(2 * np.pi * np.e * Cw)[1]              # the exact density stays Gaussian
        assert abs(float(info["entropy"]) - ent_w) < 0.08, (tc, float(info["entropy"]), ent_w)
        a, b = float(info["integral_1sigma"]), g["integral_1sigma"][i]         # b is a count of ~150 ... 2000 particles
        assert abs(a - b) < 0.08 * b + 4.0 * np.sqrt(b / 1e4), (tc, a, b)
    end = hist[12.0][1]
    assert abs(float(end["entropy"]) - 0.5 * np.log(2 * np.pi * np.e * 10) * 6) < 0.08
    for key, const, tol in (("integral_1sigma", 0.0143877, 0.02), ("integral_0.5sigma", 0.000296478, 0.01),
                            ("integral_0.1sigma", 2.07554e-8, 0.005)):
        assert abs(float(end[key]) / const - 1) < tol, (key, float(end[key]), const)


def test_student_t_diffusion_follows_the_stored_run():
    """main.py mode 'diffusion' with the Student_t latent (d = 8, nu = 2 at t = 0): the tails thin out under diffusion, so the
    variational nu = exp(dist_params) + 1 must grow the way the reference's stored run shows (2 -> 25 by t = 5), with the same
    tdvp_error decay.  Bands, not digits: the stored run has 28 more parameters than the checked-in architecture (P = 393 vs
    365) and the host chi^2 draws are unseeded in the reference (sampler.py:32).  maxStep is 5e-3: with the checked-in 1e-2
    both this path and the CPU oracle go unstable near t = 0.5 at P = 365 (explicit Heun on a stiff, heavy-tailed problem).
    Measured on a B200: dist_params within 0.17 of the stored curve, tdvp_error within 25 %."""
    from vmc_pde_b200 import sampler, var_state, evolutionEq, tdvp, stepper
    g = load("ref_diff8_student")
    np.random.seed(0)
    off = np.zeros(8)
    smp = sampler.Sampler(dim=8, numChains=30, name="Student_t", mcmc_info={"offset": off, "bound": 0.25})
    vs = var_state.VarState(smp, 8, 1, 4, network_args={"intmediate": (4,), "offset": off, "latentSpaceName": "Student_t", "dim": 8})
    assert vs.numParameters == 365 and float(vs.params["params"]["dist_params"][0]) == 0.0      # nu = exp(0) + 1 = 2 (net.py:27-36)
    eq = evolutionEq.EvolutionEquation(dim=8, name="diffusion")
    st = stepper.FixedStepper(timeStep=1e-7, mode='Heun', maxStep=5e-3, increase_fac=1.3)
    T = tdvp.TDVP()
    t, checks, hist = 0.0, [0.5, 1.0, 2.0, 3.0, 5.0], {}
    while t < 5.0 + 1e-9:
        dp, dt, info = st.step(0, T, vs.get_parameters(), evolutionEq=eq, psi=vs, nSamplesTDVP=10000, nSamplesObs=10000,
                               normFunction=norm_fun, timings=None, integrals=False)
        vs.set_parameters(dp)
        if checks and t + dt >= checks[0]:
            hist[checks.pop(0)] = (t + dt, float(vs.params["params"]["dist_params"][0]), float(T.tdvp_error), float(T.solverResidual),
                                   float(info["entropy"]))
        t += dt
    tw = g["times"]
    for tc, (tt, dpar, err, res, ent) in hist.items():
        i = int(np.argmin(np.abs(tw - tt)))
        assert abs(dpar - g["dist_params"][i][0]) < 0.3, (tc, dpar, g["dist_params"][i][0])
        assert 0.5 < err / g["tdvp_error"][i] < 2.0, (tc, err, g["tdvp_error"][i])
        assert res < 1e-5 and abs(ent - g["entropy"][i]) < 2.5
    assert hist[5.0][1] > 2.9                                   # nu > 19: the latent has become nearly Gaussian
