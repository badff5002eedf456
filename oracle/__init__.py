"""CPU oracle for the vmc_pde TDVP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain NumPy / torch float64 on the CPU, the algorithm of the
reference's TDVP time-step path (sampler.py, net.py, var_state.py, evolutionEq.py, tdvp.py,
stepper.py, mpi_wrapper.py under /root/reference/vmc_fluids).  Every function cites the
reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  Nothing under vmc_pde_b200/ imports it; the product path has no CPU fallback.

PARITY PINS.  The reference holds no tests or golden vectors, and JAX/flax are not installable in this
image, so the reference cannot be run here.  It does hold the stored OUTPUTS of its own
JAX runs (paper_plot/data_*/**/infos.hdf5); they are extracted into tests/golden/ref_*.npz by
tests/golden/make_reference_pins.py and the oracle is held to them (tests/test_reference_pins.py):

  * oracle/threefry.py + oracle/exact_dyn.py reproduce all 1201 records of the stored particle run (exact_dyn.py main
    loop: PRNGKey, split, per-particle split, float64 normal, the integrator) to round-off (1e-12) -- PINNED;
  * the sampler key chain, the multivariate_normal layout, the observables and the ball-integral draws
    (sampler.py:57-60,72-86; tdvp.py:143-162) reproduce the first stored record of the two Gauss-latent TDVP runs
    (d=8 diffusion: first right-hand side; d=6 phase space: second right-hand side of the first Heun step) to
    1e-5 ... 3e-3 where the Monte-Carlo scatter between two draws is 1e-2 ... 4e-2 -- PINNED up to the flax
    initialisation stream (hidden kernels U[-1,1) are not reproducible without flax; the last-layer kernels are 1e-5);
  * S, F, theta_dot themselves are stored nowhere by the reference (only eigenvalues, residual, tdvp_error of runs whose
    random initial network cannot be regenerated): for them the oracle stays pinned by (1) public Threefry-2x32-20 known
    answers, (2) analytic known-answer tests (depth-0 Gaussian diffusion: d/dt L_diag = D exactly), (3) the analytic
    curves the reference plots against, (4) statistical bands of the stored spectra (largest eigenvalue, number of
    eigenvalues under the cut-off, residual, tdvp_error at t=0).  PARITY UNPINNED for those quantities in the strict sense.
"""
