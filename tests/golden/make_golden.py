"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot run here: no JAX/flax in the image,
SURVEY 8c).  Run from the repository root:  python tests/golden/make_golden.py
The vectors pin the oracle against regressions and give the CUDA path a fixed target that does not need the
oracle's autograd at test time."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import flow, tdvp, threefry

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, d, depth, h, variant, latent, eqname, n, offset, seed):
    ups, downs, _ = flow.make_index_splits(d, depth, 1)
    spec = flow.FlowSpec(dim=d, depth=depth, hidden=(h,), latent=latent, variant=variant, offset=offset, inds_up=ups, inds_down=downs)
    rng = np.random.default_rng(seed)
    theta = flow.init_params(spec, seed) + 0.02 * rng.normal(size=spec.num_params)
    st = flow.OracleState(spec, theta)
    if latent == "Student_t":
        chi = np.random.default_rng(seed + 1).chisquare(float(np.exp(theta[spec.slices()[0]["dist_params"][0]]) + 1), size=n)
        st.chi2 = lambda nu, m: chi
    else:
        chi = np.zeros(0)
    key_before = st.key.copy()
    x, lp_s, z = st.sample(n)
    E, O, lp, g = tdvp.local_terms(st, x, eqname, 0.25)
    T = tdvp.OracleTDVP()
    upd = T.solve(E.numpy(), O.numpy(), lp.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), dim=d, depth=depth, hidden=h, variant=variant, latent=latent,
                        equation=eqname, t=0.25, offset=np.asarray(offset, float), inds_up=np.asarray(ups), inds_down=np.asarray(downs),
                        theta=theta, sampler_key=key_before, chi2=chi, n=n, z=z.numpy(), x=x.numpy(), logp=lp.numpy(),
                        eloc=E.numpy(), grad=g.numpy(), O_head=O.numpy()[:8], S0=T.S0, SExp=T.SExp, F0=T.F0, ev=T.ev,
                        update=upd, residual=T.solverResidual, tdvp_error=T.tdvp_error, snr=T.snr)
    print(name, "P", spec.num_params, "resid", T.solverResidual)


if __name__ == "__main__":
    case("c1_mwe", 2, 4, 1, "no_add", "Gauss", "diffusion", 512, np.zeros(2), 1)
    case("phase_space", 6, 2, 3, "different_add", "Gauss", "advection_hamiltonian_wDiss", 384, np.array([1., 0, 0, 1, 0, 0]), 2)
    case("student_t", 4, 2, 3, "no_add", "Student_t", "diffusion_drift", 256, np.zeros(4), 3)
    # RNG stream golden: first normals / uniforms of PRNGKey(0) in JAX's float64 layout
    k = threefry.prng_key(0)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), split0=threefry.split(k), normal64=threefry.normal(k, 64),
                        uniform64=threefry.uniform(k, 64), bits=threefry.random_bits(k, 64, 16))
