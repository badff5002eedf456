"""CPU: pins of the oracle (tests/ may import oracle/; the product never does)."""
import numpy as np
import pytest
import torch

from oracle import flow, tdvp, threefry


def test_threefry_random123_known_answers():
    # Random123 / jax tests/random_test.py testThreefry2x32 vectors
    def blk(k0, k1, c0, c1):
        a, b = threefry.threefry2x32(k0, k1, np.array([c0], np.uint32), np.array([c1], np.uint32))
        return int(a[0]), int(b[0])
    assert blk(0, 0, 0, 0) == (0x6b200159, 0x99ba4efe)
    assert blk(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (0x1cb996fc, 0xbb002be7)
    assert blk(0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3) == (0xc4923a9c, 0x483df7a0)


def test_jax_key_layout_pins():
    # jax.random.split(PRNGKey(0)); uniform/normal(PRNGKey(0)) float32 -- published outputs of JAX's threefry PRNG
    assert threefry.split(threefry.prng_key(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert abs(float(threefry.uniform(threefry.prng_key(0), 1, dtype=np.float32)[0]) - 0.41845703) < 1e-7
    assert abs(float(threefry.normal(threefry.prng_key(0), 1, dtype=np.float32)[0]) + 0.20584226) < 1e-6


def test_product_key_arithmetic_matches_oracle():
    from vmc_pde_b200 import _threefry as pt
    for seed in (0, 1, 12345):
        k, ko = pt.PRNGKey(seed), threefry.prng_key(seed)
        assert np.array_equal(k, ko)
        assert np.array_equal(pt.split(k, 3), threefry.split(ko, 3))
        assert np.array_equal(pt.uniform01(k, 7), threefry.uniform01(ko, 7))
        assert np.array_equal(pt.permutation(k, 12), threefry.shuffle(ko, 12))


def test_rng_golden():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rng.npz"))
    k = threefry.prng_key(0)
    assert np.array_equal(threefry.random_bits(k, 64, 16), g["bits"])
    assert np.array_equal(threefry.uniform(k, 64), g["uniform64"])
    assert np.allclose(threefry.normal(k, 64), g["normal64"], rtol=0, atol=1e-15)
    from scipy.stats import norm
    # normals are the inverse-CDF transform of the uniforms
    u = threefry.uniform(k, 64, np.nextafter(-1.0, 0.0), 1.0)
    assert np.allclose(threefry.normal(k, 64), norm.ppf((u + 1) / 2), atol=1e-12)


@pytest.mark.parametrize("d", [2, 6])
def test_kat_depth0_gaussian_diffusion(d):
    """SURVEY section 4 KAT 1: zero coupling blocks, Gauss latent, D=1, theta=0 => d/dt L_diag = D/sigma^2 = 1 exactly."""
    spec = flow.FlowSpec(dim=d, depth=0, hidden=(d // 2,))
    st = flow.OracleState(spec, np.zeros(spec.num_params))
    T = tdvp.OracleTDVP()
    upd, _ = T.rhs(st, np.zeros(spec.num_params), "diffusion", 20000, observables=False)
    sl, _ = spec.slices()
    a, b, _ = sl["L_diag"]
    assert np.allclose(upd[a:b], 1.0, atol=1e-12)
    assert np.abs(np.delete(upd, np.arange(a, b))).max() < 1e-12
    assert T.solverResidual < 1e-13 and 1.5 < T.ev[-1] < 2.6


def test_kat_reference_defaults_d2():
    """SURVEY section 4 KAT 2: depth 4, d=2, P=37: min-norm solve gives theta_dot[L_diag] ~ 1/(1+4 alpha^2)."""
    ups, downs, _ = flow.make_index_splits(2, 4, 1)
    spec = flow.FlowSpec(dim=2, depth=4, hidden=(1,), inds_up=ups, inds_down=downs)
    assert spec.num_params == 37
    th = flow.init_params(spec, 1)
    st = flow.OracleState(spec, th)
    T = tdvp.OracleTDVP()
    upd, info = T.rhs(st, th, "diffusion", 10000)
    sl, _ = spec.slices()
    a, b, _ = sl["L_diag"]
    assert np.allclose(upd[a:b], 1.0 / (1.0 + 4 * flow.ALPHA ** 2), rtol=0.05)
    assert 20 <= int((np.abs(T.ev / T.ev[-1]) < 1e-11).sum()) <= 28
    assert T.tdvp_error < 5e-3 and T.solverResidual < 1e-9
    # analytic constants the reference plots against (visualization.py:188): entropy of N(0, I_2) and unit covariance
    assert abs(info["entropy"] - 0.5 * 2 * np.log(2 * np.pi * np.e)) < 0.05
    assert np.allclose(info["covar"], np.eye(2), atol=0.05)
    assert abs(info["integral_1sigma"] - 1.0) < 0.02  # ball of radius sqrt(10) holds ~99.3% of N(0, I_2)


def test_parameter_counts_match_survey():
    """d=2->37, d=8 Gauss->364 (Student-t 365), d=6 different_add->411, d=12->762."""
    def P(d, depth, h, variant="no_add", latent="Gauss"):
        ups, downs, _ = flow.make_index_splits(d, depth, 1)
        return flow.FlowSpec(dim=d, depth=depth, hidden=(h,), variant=variant, latent=latent, inds_up=ups, inds_down=downs).num_params
    assert P(2, 4, 1) == 37 and P(8, 4, 4) == 364 and P(8, 4, 4, latent="Student_t") == 365
    assert P(6, 4, 3, "different_add") == 411 and P(12, 4, 6) == 762
    # SURVEY 8d sizes: C2 2053, C3 8187, C4 16385
    assert P(2, 4, 85) == 2053 and P(6, 8, 36, "different_add") == 8187 and P(10, 4, 185) == 16385


def test_flow_invertibility():
    ups, downs, _ = flow.make_index_splits(6, 3, 1)
    spec = flow.FlowSpec(dim=6, depth=3, hidden=(4,), variant="different_add", inds_up=ups, inds_down=downs, offset=np.arange(6) * 0.1)
    th = flow.init_params(spec, 3) + 0.05 * np.random.default_rng(0).normal(size=spec.num_params)
    st = flow.OracleState(spec, th)
    for xi in torch.randn(4, 6):
        z, lj = flow.inn(xi, st.theta, spec, st.sl, inv=False)
        xb, lji = flow.inn(z, st.theta, spec, st.sl, inv=True)
        assert torch.allclose(xb, xi, atol=1e-12) and abs(float(lj + lji)) < 1e-12
    # sampling density equals evaluation density at the sampled point (net.py:209-217)
    x, lp, _ = st.sample(16)
    assert torch.allclose(st.logp(x), lp, atol=1e-11)


def test_oracle_reproduces_goldens():
    import os
    for name in ("c1_mwe", "phase_space", "student_t"):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        spec = flow.FlowSpec(dim=int(g["dim"]), depth=int(g["depth"]), hidden=(int(g["hidden"]),), latent=str(g["latent"]),
                             variant=str(g["variant"]), offset=g["offset"], inds_up=g["inds_up"].tolist(), inds_down=g["inds_down"].tolist())
        st = flow.OracleState(spec, g["theta"])
        st.chi2 = lambda nu, m: g["chi2"]
        x, lp_s, z = st.sample(int(g["n"]))
        assert np.allclose(x.numpy(), g["x"], atol=1e-13)
        E, O, lp, gr = tdvp.local_terms(st, x, str(g["equation"]), float(g["t"]))
        assert np.allclose(E.numpy(), g["eloc"], rtol=1e-11, atol=1e-11)
        assert np.allclose(O.numpy()[:8], g["O_head"], rtol=1e-11, atol=1e-12)


def test_steppers_against_closed_form():
    """stepper.py semantics on y' = -y: Heun one step = y (1 - dt + dt^2/2) with dt grown BEFORE the step."""
    f = lambda y, k: -y
    y, dt = tdvp.heun_step(f, np.array([1.0, 2.0]), 1e-2, 1.0, 1.3)
    assert abs(dt - 1.3e-2) < 1e-15 and np.allclose(y, np.array([1.0, 2.0]) * (1 - dt + dt * dt / 2))
    y, dt = tdvp.euler_step(f, np.array([1.0]), 1e-2, 1.0, 2.0)
    assert np.allclose(y, 1 - 2e-2)
    y, rdt, ndt = tdvp.adaptive_heun_step(f, np.array([1.0]), 0.1, 1e-6, 1.0, lambda v: float(v @ v))
    assert rdt <= 0.1 and ndt <= 1.0 and abs(y[0] - np.exp(-rdt)) < 1e-4


def test_particle_oracle_statistics():
    """oracle/exact_dyn.py (exact_dyn.py:56-84): one step of the stochastic scheme reproduces drift and noise strength of
    the phase-space equation to O(dt^2) and Monte-Carlo accuracy, and is a pure function of the key."""
    from oracle import exact_dyn as ed, threefry as tf
    p = {"m": 1.0, "omega": 1.0, "lam": 0.0, "T": 10.0, "gamma": 1.0, "t": 0.0, "D": 1.0}
    rng = np.random.default_rng(0)
    N, dt = 4000, 1e-2
    x0 = np.tile(np.array([1.0, 0.5]), (N, 1))
    key = tf.prng_key(3)
    x1 = ed.integrate(x0, dt, p, ed.velocity_hamiltonian, ed.update_phase_space, key)
    assert np.array_equal(x1, ed.integrate(x0, dt, p, ed.velocity_hamiltonian, ed.update_phase_space, key))
    # mean drift: dx = p dt, dp = (-x - gamma p) dt ; noise only on p.  The scheme weights the four stage noises
    # (1, 2, 2, 1)/6 with stage steps dt/6, dt/3, dt/3, dt/6: Var[dp] = 2 m gamma T dt (1/6 + 2/3 + 2/3 + 1/6) / ... -> computed below
    assert abs(x1[:, 0].mean() - (1.0 + 0.5 * dt)) < 5e-4
    assert abs(x1[:, 1].mean() - (0.5 + (-1.0 - 0.5) * dt)) < 3 * np.sqrt(2 * 10 * dt * 1.5 / N) + 1e-3
    w = np.array([1, 2, 2, 1]) / 6.0; dts = np.array([dt / 6, dt / 3, dt / 3, dt / 6])
    var_expected = (dt ** 2) * np.sum(w ** 2 * 2 * p["m"] * p["gamma"] * p["T"] / dts)
    assert abs(x1[:, 1].var() / var_expected - 1) < 0.1
    # pure diffusion: isotropic, zero mean
    y1 = ed.integrate(np.zeros((N, 2)), dt, p, None, ed.update_diffusion, key)
    vd = (dt ** 2) * np.sum(w ** 2 * 2 / dts)
    assert abs(y1.mean()) < 4 * np.sqrt(vd / N) and abs(y1.var() / vd - 1) < 0.1
