"""CPU oracle for the vmc_pde TDVP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain NumPy / torch float64 on the CPU, the algorithm of the
reference's TDVP time-step path (sampler.py, net.py, var_state.py, evolutionEq.py, tdvp.py,
stepper.py, mpi_wrapper.py under /root/reference/vmc_fluids).  Every function cites the
reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  Nothing under vmc_pde_b200/ imports it; the product path has no CPU fallback.

PARITY UNPINNED: the reference holds no tests, golden vectors or fixtures for this path
(SURVEY.md section 8c), and JAX/flax are not installable in this image, so the reference
itself cannot be run.  The oracle is pinned instead by (1) public Threefry-2x32-20 known
answers and the well-known jax.random.split(PRNGKey(0)) / uniform(PRNGKey(0)) outputs,
(2) analytic known-answer tests (depth-0 Gaussian diffusion: d/dt L_diag = D exactly),
(3) the analytic constants the reference's plotting scripts compare against.
"""
