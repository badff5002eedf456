"""GPU parity tests, kernel level: every call goes through the C-ABI (libvmcpde.so) and is checked against the CPU
oracle on the same seeded inputs.  Tolerances: float64 round-off, relative 1e-11 unless stated (the reference path is
float64, main.py:2); RNG bits are exact, normals within a few ulp of scipy's erfinv."""
import ctypes as C
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import flow, tdvp, threefry


@pytest.fixture(scope="module")
def L():
    from vmc_pde_b200 import _lib
    _lib.require_cuda()
    return _lib.load()


def dev():
    return torch.device("cuda:0")


def relerr(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


CASES = [(2, 4, 1, "no_add", "Gauss", "diffusion", 1000),
         (2, 3, 6, "no_add", "Gauss", "advection_paper", 333),
         (6, 3, 5, "different_add", "Gauss", "advection_hamiltonian_wDiss", 777),
         (8, 4, 4, "no_add", "Student_t", "diffusion", 2048),
         (4, 2, 3, "add_s", "Gauss", "diffusion_anisotropic", 515),
         (4, 2, 3, "jac_eq_1", "Student_t", "advection_hamiltonian", 100),
         (10, 2, 20, "no_add", "Gauss", "diffusion_drift", 300),
         (12, 1, 6, "no_add", "Gauss", "diffusion_anisotropic", 64),
         (3, 2, 3, "no_add", "Gauss", "diffusion", 50),
         (5, 2, 4, "different_add", "Gauss", "diffusion_drift", 50),
         (2, 12, 2, "no_add", "Gauss", "diffusion", 40),
         (6, 0, 1, "no_add", "Gauss", "diffusion", 40),
         (6, 3, (4, 7), "different_add", "Gauss", "advection_hamiltonian_wDiss", 300),   # several hidden layers (net.py:53-58)
         (4, 2, (3, 5, 2), "no_add", "Student_t", "diffusion", 200),
         (4, 3, 3, "no_add+gc", "Gauss", "diffusion", 300),                          # global_change (net.py:72,80-82,115-116,149-150)
         (6, 2, 4, "different_add+gc", "Student_t", "advection_hamiltonian_wDiss", 200)]


@pytest.mark.parametrize("d,depth,h,variant,latent,eqname,n", CASES)
def test_sampler_local_terms_moments_gram(L, d, depth, h, variant, latent, eqname, n):
    from vmc_pde_b200 import _lib, _capi
    rng = np.random.default_rng(d * 1000 + depth * 10 + n)
    ups, downs, _ = flow.make_index_splits(d, depth, 1)
    off = rng.normal(size=d) * 0.3
    hidden = h if isinstance(h, tuple) else (h,)
    gc = variant.endswith("+gc")
    variant = variant.replace("+gc", "")
    spec = flow.FlowSpec(dim=d, depth=depth, hidden=hidden, latent=latent, variant=variant, offset=off, inds_up=ups, inds_down=downs,
                         global_change=gc)
    th = flow.init_params(spec, 1) + 0.05 * rng.normal(size=spec.num_params)
    for name, (a_, b_, _) in spec.slices()[0].items():
        if name.endswith("global_scale"):
            th[a_:b_] = rng.uniform(0.7, 1.4)
        if name.endswith("global_offset"):
            th[a_:b_] = 0.3 * rng.normal(size=b_ - a_)
    st = flow.OracleState(spec, th)
    cfg, keep = _capi.make_flow_config(d, depth, hidden, variant, latent, ups, downs, off, global_change=gc)
    fh = C.c_void_p()
    _lib.check(L.vmcpde_flow_create(C.byref(cfg), C.byref(fh)))
    P = L.vmcpde_flow_num_params(fh)
    assert P == spec.num_params
    Pp = L.vmcpde_padded_params(P)
    f64 = torch.float64
    tht = torch.tensor(th, device=dev())
    # ---- sampler: bit-exact counters, same latent draw, same flow inverse
    x = torch.empty(n, d, device=dev(), dtype=f64); lp = torch.empty(n, device=dev(), dtype=f64); z = torch.empty_like(x)
    chi2 = None
    if latent == "Student_t":
        chi2_np = np.random.default_rng(5).chisquare(float(np.exp(th[spec.slices()[0]["dist_params"][0]]) + 1), size=n)
        chi2 = torch.tensor(chi2_np, device=dev()); st.chi2 = lambda nu, m: chi2_np
    key = threefry.split(st.key, 2)[1]
    _lib.check(L.vmcpde_sample(fh, _lib.ptr(tht), int(key[0]), int(key[1]), 0, n, n, _lib.ptr(chi2), _lib.ptr(x), _lib.ptr(lp), _lib.ptr(z), _lib.stream()))
    xo, lpo, zo = st.sample(n)
    assert relerr(z, zo) < 1e-13 and relerr(x, xo) < 1e-12 and relerr(lp, lpo) < 1e-12
    # a shard of the same stream reproduces the corresponding slice exactly
    a, m = n // 3, n // 2
    xs = torch.empty(m, d, device=dev(), dtype=f64); lps = torch.empty(m, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_sample(fh, _lib.ptr(tht), int(key[0]), int(key[1]), a, m, n, _lib.ptr(chi2[a:a + m].contiguous() if chi2 is not None else None),
                               _lib.ptr(xs), _lib.ptr(lps), None, _lib.stream()))
    assert torch.equal(xs, x[a:a + m]) and torch.equal(lps, lp[a:a + m])
    # ---- fused local terms on the oracle's samples
    xin = torch.tensor(xo.numpy(), device=dev())
    A = torch.tensor(tdvp.random_D_factor(d), device=dev()) if eqname == "diffusion_anisotropic" else None
    eq = _capi.make_equation(eqname, dict(tdvp.EQ_PARAMS.get(eqname, {})), 0.3, A.data_ptr() if A is not None else None)
    nrow = (n + 15) // 16 * 16
    E = torch.empty(n, device=dev(), dtype=f64); lp2 = torch.empty_like(E); g = torch.empty(n, d, device=dev(), dtype=f64); lap = torch.empty_like(E)
    O = torch.full((nrow, Pp), float("nan"), device=dev(), dtype=f64)
    O[n:] = 0
    _lib.check(L.vmcpde_local_terms(fh, _lib.ptr(tht), _lib.ptr(xin), n, C.byref(eq), _lib.ptr(E), _lib.ptr(lp2), _lib.ptr(g), _lib.ptr(lap), _lib.ptr(O), Pp, _lib.stream()))
    Eo, Oo, lpo2, go = tdvp.local_terms(st, xo, eqname, 0.3)
    assert relerr(E, Eo) < 1e-11 and relerr(O[:n, :P], Oo) < 1e-11 and relerr(lp2, lpo2) < 1e-12
    assert not torch.isnan(O).any() and (Pp == P or float(O[:, P:].abs().max()) == 0.0)   # every column written, padding zeroed
    if eqname != "diffusion_anisotropic":
        assert relerr(g, go) < 1e-11
    lp3 = torch.empty_like(E)
    _lib.check(L.vmcpde_logp(fh, _lib.ptr(tht), _lib.ptr(xin), n, _lib.ptr(lp3), _lib.stream()))
    assert relerr(lp3, lpo2) < 1e-12
    nh = min(n, 48)
    H = torch.empty(nh, d, d, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_hessian(fh, _lib.ptr(tht), _lib.ptr(xin), nh, _lib.ptr(H), _lib.stream()))
    assert relerr(H, st.hessian(xo[:nh])) < 1e-10
    # flow transform round trip (main.py:77-96)
    y = torch.empty_like(xin); lj = torch.empty_like(E); xb = torch.empty_like(xin); lji = torch.empty_like(E)
    _lib.check(L.vmcpde_flow_transform(fh, _lib.ptr(tht), _lib.ptr(xin), n, 0, _lib.ptr(y), _lib.ptr(lj), None, _lib.stream()))
    _lib.check(L.vmcpde_flow_transform(fh, _lib.ptr(tht), _lib.ptr(y), n, 1, _lib.ptr(xb), _lib.ptr(lji), None, _lib.stream()))
    if not gc:
        assert relerr(xb, xin) < 1e-11 and float((lj + lji).abs().max()) < 1e-10
    else:   # the reference's inverse branch is not the inverse of its forward branch (net.py:115-116 vs 149-150); the log-Jacobians still cancel
        assert relerr(xb, xin) > 1e-3
    # ---- first moments, centring, force, three weighted Grams (tdvp.py:36-52, 68-70)
    T = tdvp.OracleTDVP(); T.solve(Eo.numpy(), Oo.numpy(), lpo2.numpy())
    sums = torch.zeros(4 + Pp, device=dev(), dtype=f64)
    mws = torch.empty(((n + 511) // 512 + 1) * Pp, device=O.device, dtype=torch.float64)
    _lib.check(L.vmcpde_moments1(_lib.ptr(E), _lib.ptr(lp2), _lib.ptr(O), n, Pp, _lib.ptr(sums), _lib.ptr(mws), mws.numel() * 8, _lib.stream()))
    assert abs(float(sums[0]) / n - T.ElocMean) <= 1e-11 * (abs(T.ElocMean) + np.abs(Eo.numpy()).max())
    assert abs(float(sums[1]) / n - T.ElocMeanAbs) <= 1e-11 * T.ElocMeanAbs
    assert relerr(sums[4:4 + P] / n, T.gradMean) < 1e-10
    meanO = (sums[4:] / n).contiguous()
    dE = torch.zeros(nrow, device=dev(), dtype=f64); wE = torch.zeros_like(dE); wLp = torch.zeros_like(dE)
    F = torch.zeros(Pp, device=dev(), dtype=f64); var = torch.zeros(1, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_center_force(_lib.ptr(O), n, Pp, _lib.ptr(meanO), _lib.ptr(E), _lib.ptr(lp2), float(sums[0]) / n, _lib.ptr(dE), _lib.ptr(wE), _lib.ptr(wLp), _lib.ptr(F), _lib.ptr(var), _lib.ptr(mws), mws.numel() * 8, _lib.stream()))
    assert relerr(F[:P] / n, T.F0) < 1e-10 and abs(float(var) / n / T.ElocVar - 1) < 1e-11
    S = [torch.zeros(Pp, Pp, device=dev(), dtype=f64) for _ in range(3)]
    _lib.check(L.vmcpde_gram(_lib.ptr(O), nrow, Pp, Pp, 3, _lib.ptr_array([None, wLp, wE]), _lib.ptr_array(S), _lib.stream()))
    for s_ in S:
        _lib.check(L.vmcpde_sym_finalize(_lib.ptr(s_), Pp, 1.0 / n, _lib.stream()))
    dO = Oo.numpy() - T.gradMean; dEo = Eo.numpy() - T.ElocMean
    Ceo = (dO * (dEo ** 2)[:, None]).T @ dO / n
    assert relerr(S[0][:P, :P], T.S0) < 1e-11 and relerr(S[1][:P, :P], T.SExp) < 1e-11 and relerr(S[2][:P, :P], Ceo) < 1e-11
    assert float((S[0] - S[0].T).abs().max()) == 0.0
    L.vmcpde_flow_destroy(fh)


def test_rng_streams_bit_exact_and_golden(L):
    from vmc_pde_b200 import _lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rng.npz"))
    k = threefry.prng_key(0)
    u = torch.empty(64, device=dev(), dtype=torch.float64); nrm = torch.empty_like(u)
    _lib.check(L.vmcpde_uniform(int(k[0]), int(k[1]), 0, 64, 64, _lib.ptr(u), _lib.stream()))
    _lib.check(L.vmcpde_normal(int(k[0]), int(k[1]), 0, 64, 64, _lib.ptr(nrm), _lib.stream()))
    assert np.array_equal(u.cpu().numpy(), g["uniform64"])                       # integer path + mantissa trick: exact
    assert np.allclose(nrm.cpu().numpy(), g["normal64"], rtol=1e-14, atol=1e-15)  # erfinv: a few ulp
    # offsets into a longer stream
    k2 = threefry.prng_key(7)
    part = torch.empty(100, device=dev(), dtype=torch.float64)
    _lib.check(L.vmcpde_uniform(int(k2[0]), int(k2[1]), 900, 100, 5000, _lib.ptr(part), _lib.stream()))
    assert np.array_equal(part.cpu().numpy(), threefry.uniform(k2, 5000)[900:1000])
    # range validation
    assert L.vmcpde_uniform(0, 0, 10, 100, 50, _lib.ptr(part), _lib.stream()) != 0
    assert L.vmcpde_normal(0, 0, 0, 1, 2 ** 31 + 1, _lib.ptr(part), _lib.stream()) != 0  # 32-bit counter range


def test_gram_properties_large(L):
    """Size-independent properties at a BASELINE-sized panel (P = 8192): agreement with an independent FP64 product on a
    slab, symmetry, linearity in the row weights, additivity over sample chunks, empty input."""
    from vmc_pde_b200 import _lib
    n, Pp = 4096, 8192
    f64 = torch.float64
    O = torch.randn(n, Pp, device=dev(), dtype=f64)
    w = torch.rand(n, device=dev(), dtype=f64)
    S = [torch.zeros(Pp, Pp, device=dev(), dtype=f64) for _ in range(2)]
    _lib.check(L.vmcpde_gram(_lib.ptr(O), n, Pp, Pp, 2, _lib.ptr_array([None, w]), _lib.ptr_array(S), _lib.stream()))
    ref = O[:, 1000:1256].T @ O
    refw = (O[:, 1000:1256] * w[:, None]).T @ O
    up = torch.triu(torch.ones(Pp, Pp, device=dev(), dtype=torch.bool))[1000:1256]
    assert float(((S[0][1000:1256] - ref) * up).abs().max()) < 1e-10 * float(ref.abs().max())
    assert float(((S[1][1000:1256] - refw) * up).abs().max()) < 1e-10 * float(refw.abs().max())
    # two chunks accumulate to the same result; doubling the weights doubles the matrix
    S2 = torch.zeros(Pp, Pp, device=dev(), dtype=f64)
    w2 = (2 * w).contiguous()
    for c0 in (0, n // 2):
        _lib.check(L.vmcpde_gram(_lib.ptr(O[c0:c0 + n // 2]), n // 2, Pp, Pp, 1, _lib.ptr_array([w2[c0:c0 + n // 2]]), _lib.ptr_array([S2]), _lib.stream()))
    assert float((torch.triu(S2) - 2 * torch.triu(S[1])).abs().max()) < 1e-9 * float(S[1].abs().max())
    before = S2.clone()
    _lib.check(L.vmcpde_gram(_lib.ptr(O), 0, Pp, Pp, 1, _lib.ptr_array([None]), _lib.ptr_array([S2]), _lib.stream()))
    assert torch.equal(before, S2)
    assert L.vmcpde_gram(_lib.ptr(O), 17, Pp, Pp, 1, _lib.ptr_array([None]), _lib.ptr_array([S2]), _lib.stream()) != 0  # n % 16
    _lib.check(L.vmcpde_sym_finalize(_lib.ptr(S[0]), Pp, 1.0, _lib.stream()))
    assert torch.equal(S[0], S[0].T)


def test_gram_tail_split_syrk_and_splitk_products(L):
    """The partly filled last round of work items is split along the samples (513 items on 148 CTAs: 69 tail items, 2
    slices each, added in slice order): every tile against an independent FP64 product, bit-for-bit repeatable.  Plus the
    two other forms of the pipeline: vmcpde_syrk_tn (scaled, upper tiles only) and vmcpde_gemm_tn_splitk."""
    from vmc_pde_b200 import _lib
    f64 = torch.float64
    n, Pp = 8192, 2304
    O = torch.randn(n, Pp, device=dev(), dtype=f64)
    w1, w2 = torch.rand(n, device=dev(), dtype=f64), torch.rand(n, device=dev(), dtype=f64)
    runs = []
    for _ in range(2):
        S = [torch.zeros(Pp, Pp, device=dev(), dtype=f64) for _ in range(3)]
        _lib.check(L.vmcpde_gram(_lib.ptr(O), n, Pp, Pp, 3, _lib.ptr_array([None, w1, w2]), _lib.ptr_array(S), _lib.stream()))
        runs.append(S)
    up = torch.triu(torch.ones(Pp, Pp, device=dev(), dtype=torch.bool))
    blockup = up.view(Pp // 128, 128, Pp // 128, 128).any(dim=3).any(dim=1)                      # computed tiles
    tilemask = blockup.repeat_interleave(128, 0).repeat_interleave(128, 1)
    for S, w in zip(runs[0], (None, w1, w2)):
        ref = (O if w is None else O * w[:, None]).T @ O
        assert float(((S - ref) * tilemask).abs().max()) < 1e-11 * float(ref.abs().max())
    assert all(torch.equal(a, b) for a, b in zip(runs[0], runs[1]))
    # SYRK form with a scale: Out(upper tiles) = alpha X^T X + beta Out
    K, M = 256, 640
    X = torch.randn(K, M, device=dev(), dtype=f64); Out = torch.randn(M, M, device=dev(), dtype=f64); Out0 = Out.clone()
    _lib.check(L.vmcpde_syrk_tn(_lib.ptr(X), M, _lib.ptr(Out), M, M, K, -0.5, 2.0, _lib.stream()))
    ref = -0.5 * X.T @ X + 2.0 * Out0
    tm = torch.triu(torch.ones(M // 128, M // 128, device=dev(), dtype=torch.bool)).repeat_interleave(128, 0).repeat_interleave(128, 1)
    assert float(((Out - ref) * tm).abs().max()) < 1e-12 * float(ref.abs().max()) and torch.equal(Out[~tm], Out0[~tm])
    # split-K: slices of the contraction in separate outputs
    K, M, N, sp = 1040, 128, 256, 5
    X = torch.randn(K, M, device=dev(), dtype=f64); Y = torch.randn(K, N, device=dev(), dtype=f64)
    Part = torch.full((sp, M, N), float("nan"), device=dev(), dtype=f64)
    _lib.check(L.vmcpde_gemm_tn_splitk(_lib.ptr(X), M, _lib.ptr(Y), N, _lib.ptr(Part), N, M, N, K, sp, _lib.stream()))
    assert relerr(Part.sum(0), X.T @ Y) < 1e-13
    kper = -(-(K // 16) // sp) * 16
    assert relerr(Part[1], X[kper:2 * kper].T @ Y[kper:2 * kper]) < 1e-13


def test_gemm_tn(L):
    from vmc_pde_b200 import _lib
    K, M, N = 528, 256, 384
    X = torch.randn(K, M, device=dev(), dtype=torch.float64); Y = torch.randn(K, N, device=dev(), dtype=torch.float64)
    Out = torch.randn(M, N, device=dev(), dtype=torch.float64)
    ref = 0.5 * X.T @ Y + 2.0 * Out
    _lib.check(L.vmcpde_gemm_tn(_lib.ptr(X), M, _lib.ptr(Y), N, _lib.ptr(Out), N, M, N, K, 0.5, 2.0, _lib.stream()))
    assert relerr(Out, ref) < 1e-13


@pytest.mark.parametrize("n,P", [(1000, 250), (256, 128), (77, 384), (5000, 640)])
def test_gram_split_tcgen05_within_the_stated_tolerance(L, n, P):
    """vmcpde_gram_split (bf16 x 3 split operands on tcgen05, FP32 TMEM accumulation over <= 256 samples, FP64 sums) against
    the FP64 Gram: |error| <= 1e-6 sqrt(S_ii S_jj) entry by entry -- the tolerance stated in include/vmcpde.h -- on a graded
    matrix (6 decades of column scale), ragged sample counts, with and without row weights; accumulates into S."""
    from vmc_pde_b200 import _lib
    rng = np.random.default_rng(n + P)
    Pp = L.vmcpde_padded_params(P)
    O_np = np.zeros((n, Pp)); O_np[:, :P] = rng.normal(size=(n, P)) * 10.0 ** (-6.0 * np.arange(P) / P)
    w_np = rng.normal(size=n) ** 2 * 3.0
    O = torch.tensor(O_np, device=dev()); w = torch.tensor(w_np, device=dev())
    nb = C.c_size_t(0); _lib.check(L.vmcpde_gram_split_workspace_bytes(n, Pp, C.byref(nb)))
    ws = torch.empty(nb.value, device=dev(), dtype=torch.uint8)
    for weights, wn in ((None, np.ones(n)), (w, w_np)):
        S = torch.full((Pp, Pp), 0.5, device=dev(), dtype=torch.float64)
        _lib.check(L.vmcpde_gram_split(_lib.ptr(O), n, Pp, Pp, _lib.ptr(weights), _lib.ptr(S), _lib.ptr(ws), nb.value, _lib.stream()))
        ref = (O_np * wn[:, None]).T @ O_np
        got = S.cpu().numpy() - 0.5
        d = np.sqrt(np.diag(ref)[:P])
        iu = np.triu_indices(P)
        err = np.abs(got[:P, :P] - ref[:P, :P])[iu] / np.outer(d, d)[iu]
        assert err.max() < 1e-6, err.max()
        if Pp > P:
            assert np.abs(np.triu(got)[:, P:]).max() == 0.0      # zero padding columns stay zero


def test_packed_tiles_and_nccl_entry_points_single_rank(L):
    """vmcpde_pack/unpack_upper_tiles round trip (only the upper 128 x 128 tiles travel) and the NCCL entry points on a
    one-rank communicator created through the C-ABI itself (sum over one rank = identity; 2 ranks: tests/test_gpu_multi.py)."""
    from vmc_pde_b200 import _lib
    f64 = torch.float64
    Pp = 384
    S = torch.randn(Pp, Pp, device=dev(), dtype=f64)
    ln = L.vmcpde_packed_tiles_len(Pp)
    assert ln == 6 * 128 * 128
    packed = torch.zeros(ln, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_pack_upper_tiles(_lib.ptr(S), Pp, _lib.ptr(packed), _lib.stream()))
    back = torch.full((Pp, Pp), -1.0, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_unpack_upper_tiles(_lib.ptr(packed), Pp, _lib.ptr(back), _lib.stream()))
    for ti in range(3):
        for tj in range(3):
            blk = (slice(ti * 128, ti * 128 + 128), slice(tj * 128, tj * 128 + 128))
            assert torch.equal(back[blk], S[blk]) if tj >= ti else bool((back[blk] == -1.0).all())
    ident = C.create_string_buffer(128)
    _lib.check(L.vmcpde_nccl_unique_id(ident))
    comm = C.c_void_p()
    _lib.check(L.vmcpde_nccl_comm_init(1, 0, ident, C.byref(comm)))
    v = torch.arange(1000, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_allreduce_sum(comm, _lib.ptr(v), 1000, _lib.stream()))
    mats = [torch.randn(Pp, Pp, device=dev(), dtype=f64) for _ in range(2)]
    ref = [m.clone() for m in mats]
    tail = torch.randn(Pp + 8, device=dev(), dtype=f64); tail_ref = tail.clone()
    pk = torch.empty(2 * ln + Pp + 8, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_allreduce_moments(comm, _lib.ptr_array(mats), 2, Pp, _lib.ptr(tail), Pp + 8, _lib.ptr(pk), _lib.stream()))
    torch.cuda.synchronize()
    assert torch.equal(v, torch.arange(1000, device=dev(), dtype=f64)) and torch.equal(tail, tail_ref)
    assert all(torch.equal(torch.triu(a), torch.triu(b)) for a, b in zip(mats, ref))
    _lib.check(L.vmcpde_nccl_comm_destroy(comm))


def _eigh(L, S_np):
    from vmc_pde_b200 import _lib
    n = S_np.shape[0]; ld = L.vmcpde_padded_params(n)
    S = torch.zeros(ld, ld, device=dev(), dtype=torch.float64); S[:n, :n] = torch.tensor(S_np, device=dev())
    ev = torch.zeros(ld, device=dev(), dtype=torch.float64); VT = torch.zeros(ld, ld, device=dev(), dtype=torch.float64)
    nb = C.c_size_t(0); _lib.check(L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(nb)))
    ws = torch.empty(nb.value, device=dev(), dtype=torch.uint8)
    _lib.check(L.vmcpde_eigh(_lib.ptr(S), n, ld, _lib.ptr(ev), _lib.ptr(VT), _lib.ptr(ws), nb.value, _lib.stream()))
    return ev[:n].cpu().numpy(), VT[:n, :n].cpu().numpy().T


def test_eigh_against_lapack(L):
    rng = np.random.default_rng(0)
    mats = []
    for n in (1, 2, 3, 37, 130, 300, 385, 777, 1000):   # n >= 384 takes the blocked (cooperative panel) path
        A = rng.normal(size=(n, n)); mats.append((A + A.T) / 2)
    A = rng.normal(size=(3000, 200)) @ rng.normal(size=(200, 600)); mats.append(A.T @ A / 3000)        # rank deficient
    cs = 10.0 ** (-6.0 * np.arange(1100) / 1100); A = rng.normal(size=(3000, 1100)) * cs; mats.append(A.T @ A / 3000)  # graded
    mats.append(np.diag(rng.normal(size=50)))                                                          # already diagonal
    mats.append(np.zeros((20, 20)))
    for S in mats:
        n = S.shape[0]
        ev, V = _eigh(L, S)
        ref = np.linalg.eigvalsh(S); nrm = max(np.abs(ref).max(), 1e-300)
        assert np.all(np.diff(ev) >= 0)
        assert np.abs(ev - ref).max() <= 1e-13 * nrm + 1e-300
        assert np.abs(S @ V - V * ev).max() <= 2e-13 * nrm + 1e-300
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12


def test_eigh_blocked_path_edge_cases(L):
    """Blocked path (n >= 384): reflector-free columns (diagonal / zero / block-diagonal input), the size thresholds of the
    two mat-vec modes, sizes that are not multiples of the 4-row ownership chunks or of the 64-column panels, a spectrum
    graded over 24 decades, tiny and huge overall scales."""
    rng = np.random.default_rng(7)

    def check(S, tol_res=3e-13):
        n = S.shape[0]
        ev, V = _eigh(L, S)
        ref = np.linalg.eigvalsh(S); nrm = max(np.abs(ref).max(), 1e-300)
        assert np.all(np.isfinite(ev)) and np.all(np.isfinite(V)) and np.all(np.diff(ev) >= 0)
        assert np.abs(ev - ref).max() <= 2e-13 * nrm + 1e-300
        assert np.abs(S @ V - V * ev).max() <= tol_res * nrm + 1e-300
        assert np.abs(V.T @ V - np.eye(n)).max() < 2e-12

    check(np.diag(rng.normal(size=600)))                      # every column has sigma == 0 (tau = 0)
    check(np.zeros((400, 400)))
    check(np.eye(450) * 3.0)
    B = np.zeros((700, 700)); A = rng.normal(size=(300, 300)); B[:300, :300] = A + A.T; B[300:, 300:] = np.diag(rng.normal(size=400))
    check(B)                                                  # dense block followed by reflector-free columns
    for n in (384, 447, 449, 513, 1535, 1537, 1601):          # panel / chunk / mode boundaries (sym mode from a trailing size of 1536)
        A = rng.normal(size=(n, n)); check((A + A.T) / 2)
    n = 900
    cs = 10.0 ** (-12.0 * np.arange(n) / n); A = rng.normal(size=(2 * n, n)) * cs; check(A.T @ A / (2 * n))   # 24 decades
    A = rng.normal(size=(500, 500)); S = (A + A.T) / 2
    check(S * 1e-150); check(S * 1e150)


@pytest.mark.parametrize("n", [8187, 16385])
def test_eigh_at_benchmark_sizes_against_the_library_solver(L, n):
    """The sizes the benchmarks run (C3: P = 8187, C4: P = 16385) on a graded, numerically rank-deficient Gram matrix like
    the S of a real run: every eigenvalue against cuSOLVER's syevd (torch.linalg.eigvalsh -- yardstick only, never on the
    product path) to 1e-13 ||S||, the full orthogonality defect ||V^T V - I||_max and the full residual ||S V - V diag(ev)||_max."""
    from vmc_pde_b200 import _lib
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    rows = n // 2                                              # rank <= n/2: half of the spectrum is round-off, as in real runs
    A = torch.randn(rows, n, device=dev(), dtype=torch.float64, generator=g)
    A *= 10.0 ** (-6.0 * torch.arange(n, device=dev(), dtype=torch.float64) / n)
    ld = L.vmcpde_padded_params(n)
    S = torch.zeros(ld, ld, device=dev(), dtype=torch.float64)
    S[:n, :n] = A.T @ A / rows
    del A
    S[:n, :n] = (S[:n, :n] + S[:n, :n].T) / 2
    ref = torch.linalg.eigvalsh(S[:n, :n])
    work = S.clone()
    ev = torch.zeros(ld, device=dev(), dtype=torch.float64); VT = torch.zeros(ld, ld, device=dev(), dtype=torch.float64)
    nb = C.c_size_t(0); _lib.check(L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(nb)))
    ws = torch.empty(nb.value, device=dev(), dtype=torch.uint8)
    _lib.check(L.vmcpde_eigh(_lib.ptr(work), n, ld, _lib.ptr(ev), _lib.ptr(VT), _lib.ptr(ws), nb.value, _lib.stream()))
    del work, ws
    nrm = float(ref.abs().max())
    assert bool((ev[1:n] >= ev[:n - 1]).all())
    assert float((ev[:n] - ref).abs().max()) <= 1e-13 * nrm
    V = VT[:n, :n].T
    G = VT[:n, :n] @ V
    G.diagonal().sub_(1.0)
    assert float(G.abs().max()) < 5e-12
    del G
    R = S[:n, :n] @ V - V * ev[:n]
    assert float(R.abs().max()) <= 5e-13 * nrm
    assert int((ev[:n].abs() < 1e-11 * nrm).sum()) >= n // 2 - 8   # the null space is resolved as such


def test_eigh_two_call_form_equals_the_single_call(L):
    """vmcpde_eigh_factor + vmcpde_eigh_backtransform (the seam of the pipelined multi-GPU solve) == vmcpde_eigh bit for bit,
    slice by slice, also when the caller's Z^T / VT buffers hold garbage on entry; the unblocked sizes refuse."""
    from vmc_pde_b200 import _lib
    rng = np.random.default_rng(21)
    f64 = torch.float64
    for n in (700, 2053):
        A = rng.normal(size=(n + 50, n)) * 10.0 ** (-5.0 * np.arange(n) / n); S_np = A.T @ A / (n + 50)
        ev0, V0 = _eigh(L, S_np)
        ld = L.vmcpde_padded_params(n)
        S = torch.zeros(ld, ld, device=dev(), dtype=f64); S[:n, :n] = torch.tensor(S_np, device=dev())
        ev = torch.zeros(ld, device=dev(), dtype=f64); tau = torch.zeros(ld, device=dev(), dtype=f64)
        ZT = torch.full((ld, ld), float("nan"), device=dev(), dtype=f64)
        VT = torch.full((ld, ld), 7.0, device=dev(), dtype=f64)
        nb = C.c_size_t(0); _lib.check(L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(nb)))
        ws = torch.empty(nb.value, device=dev(), dtype=torch.uint8)
        _lib.check(L.vmcpde_eigh_factor(_lib.ptr(S), n, ld, _lib.ptr(ev), _lib.ptr(ZT), _lib.ptr(tau), _lib.ptr(ws), nb.value, _lib.stream()))
        ws.fill_(255)                                            # the factors live in caller buffers, not in the workspace
        for col0, ncols in ((0, 256), (256, ld - 256)):
            _lib.check(L.vmcpde_eigh_backtransform(_lib.ptr(S), _lib.ptr(tau), _lib.ptr(ZT), n, ld, _lib.ptr(VT), col0, ncols,
                                                   _lib.ptr(ws), nb.value, _lib.stream()))
        assert np.array_equal(ev[:n].cpu().numpy(), ev0) and np.array_equal(VT[:n, :n].cpu().numpy().T, V0)
    S = torch.eye(200, device=dev(), dtype=f64)
    assert L.vmcpde_eigh_factor(_lib.ptr(S), 200, 200, _lib.ptr(ev), _lib.ptr(ZT), _lib.ptr(tau), _lib.ptr(ws), nb.value, _lib.stream()) == 2


def test_eigh_is_run_to_run_deterministic(L):
    """Fixed-order reductions and tagged messages: two runs on the same matrix agree bit for bit (blocked path, both
    mat-vec modes), and so do the replicated solves of a multi-GPU run."""
    rng = np.random.default_rng(11)
    for n in (700, 2053):
        A = rng.normal(size=(n + 100, n)) * 10.0 ** (-4.0 * np.arange(n) / n); S = A.T @ A / (n + 100)
        runs = [_eigh(L, S) for _ in range(3)]
        for ev, V in runs[1:]:
            assert np.array_equal(ev, runs[0][0]) and np.array_equal(V, runs[0][1])


def test_solve_tail_and_cholesky_against_oracle(L):
    from vmc_pde_b200 import _lib
    rng = np.random.default_rng(3)
    f64 = torch.float64
    for (n, Ns, useSNR) in ((37, 2000, 0), (300, 4000, 1)):
        ld = L.vmcpde_padded_params(n)
        O_ = rng.normal(size=(Ns, n)) * 10.0 ** (-3.0 * np.arange(n) / n)
        O_[:, n // 2:] = O_[:, : n - n // 2] @ rng.normal(size=(n - n // 2, n - n // 2)) * 1e-1   # exactly rank deficient
        E = rng.normal(size=Ns) + 0.3; lp = rng.normal(size=Ns)
        T = tdvp.OracleTDVP(useSNR=bool(useSNR)); upd = T.solve(E, O_, lp)
        dO = O_ - O_.mean(0); dE = E - E.mean(); CEO = (dO * (dE ** 2)[:, None]).T @ dO / Ns
        def Pd(a):
            t = torch.zeros(ld, ld, device=dev(), dtype=f64); t[:n, :n] = torch.tensor(a, device=dev()); return t
        S, S0, Cg = Pd(T.S), Pd(T.S0), Pd(CEO)
        F = torch.zeros(ld, device=dev(), dtype=f64); F[:n] = torch.tensor(T.F0, device=dev())
        ev = torch.zeros(ld, device=dev(), dtype=f64); VT = torch.zeros(ld, ld, device=dev(), dtype=f64)
        a, b = C.c_size_t(0), C.c_size_t(0)
        L.vmcpde_eigh_workspace_bytes(n, ld, C.byref(a)); L.vmcpde_solve_tail_workspace_bytes(n, ld, C.byref(b))
        ws = torch.empty(max(a.value, b.value), device=dev(), dtype=torch.uint8)
        A = S.clone()
        _lib.check(L.vmcpde_eigh(_lib.ptr(A), n, ld, _lib.ptr(ev), _lib.ptr(VT), _lib.ptr(ws), ws.numel(), _lib.stream()))
        outs = [torch.zeros(ld, device=dev(), dtype=f64) for _ in range(5)]; sc = torch.zeros(2, device=dev(), dtype=f64)
        _lib.check(L.vmcpde_solve_tail(_lib.ptr(ev), _lib.ptr(VT), n, ld, _lib.ptr(F), _lib.ptr(S), _lib.ptr(S0), _lib.ptr(Cg), float(Ns), 1e-11, 2.0, useSNR,
                                       float(np.mean(E ** 2)), *[_lib.ptr(o) for o in outs], _lib.ptr(sc), _lib.ptr(ws), ws.numel(), _lib.stream()))
        VtF, rhoVar, snr, invEv, update = [o[:n].cpu().numpy() for o in outs]
        du = update - upd
        # theta_dot is compared in the S-norm: null-space components are undefined
        assert du @ T.S0 @ du <= 1e-16 * (upd @ T.S0 @ upd)
        assert relerr(ev[:n], T.ev) < 1e-13
        assert abs(float(sc[1]) - T.tdvp_error) < 1e-11 and float(sc[0]) < 10 * T.solverResidual + 1e-12
        big = np.abs(T.ev / T.ev[-1]) > 1e-8
        assert np.abs(snr[big] / T.snr[big] - 1).max() < 1e-6 and relerr(rhoVar[big], T.rhoVar[big]) < 1e-9
        assert np.array_equal(invEv == 0, T.invEv == 0)
    for n in (5, 64, 65, 300, 1000, 1281, 2053):   # n >= 256 with a padded ld takes the tensor-core (upper-form) path
        ld = L.vmcpde_padded_params(n); A = rng.normal(size=(n + 50, n)); Sn = A.T @ A / n + 1e-3 * np.eye(n); Fn = rng.normal(size=n)
        S = torch.zeros(ld, ld, device=dev(), dtype=f64); S[:n, :n] = torch.tensor(Sn, device=dev()); F = torch.tensor(Fn, device=dev())
        x = torch.zeros(n, device=dev(), dtype=f64); info = torch.zeros(1, device=dev(), dtype=torch.int32)
        _lib.check(L.vmcpde_chol_solve(_lib.ptr(S), n, ld, _lib.ptr(F), _lib.ptr(x), _lib.ptr(info), _lib.stream()))
        assert int(info) == 0 and relerr(x, np.linalg.solve(Sn, Fn)) < 1e-10
    # a singular matrix is reported, not silently "solved"
    S = torch.zeros(128, 128, device=dev(), dtype=f64); S[:4, :4] = torch.tensor(np.ones((4, 4)), device=dev())
    info = torch.zeros(1, device=dev(), dtype=torch.int32); x = torch.zeros(4, device=dev(), dtype=f64)
    _lib.check(L.vmcpde_chol_solve(_lib.ptr(S), 4, 128, _lib.ptr(torch.ones(4, device=dev(), dtype=f64)), _lib.ptr(x), _lib.ptr(info), _lib.stream()))
    assert int(info) == 2
    # ... while a parameter nothing depends on (exactly zero row and column) just gets a zero update
    n = 300; ld = L.vmcpde_padded_params(n); A = rng.normal(size=(n + 50, n)); A[:, 17] = 0.0; Sn = A.T @ A / n + 1e-3 * np.diag(np.diag(A.T @ A / n))
    Fn = rng.normal(size=n); Fn[17] = 0.0
    S = torch.zeros(ld, ld, device=dev(), dtype=f64); S[:n, :n] = torch.tensor(Sn, device=dev()); F = torch.tensor(Fn, device=dev())
    x = torch.zeros(n, device=dev(), dtype=f64); info = torch.zeros(1, device=dev(), dtype=torch.int32)
    _lib.check(L.vmcpde_chol_solve(_lib.ptr(S), n, ld, _lib.ptr(F), _lib.ptr(x), _lib.ptr(info), _lib.stream()))
    keep = np.arange(n) != 17
    assert int(info) == 0 and float(x[17]) == 0.0 and relerr(x.cpu().numpy()[keep], np.linalg.solve(Sn[np.ix_(keep, keep)], Fn[keep])) < 1e-10
