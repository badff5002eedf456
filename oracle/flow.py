"""Oracle restatement of the reference ansatz (net.py) and variational state (var_state.py).

torch float64 on the CPU; derivatives come from torch.func automatic differentiation, i.e.
they are independent of the hand-derived forward-Laplacian / reverse sweep in the CUDA kernels.

Follows: net.py:10-20 (Gauss), :23-36 (Student_t), :44-61 (SingleTrafo), :65-153 (SingleBlock),
:156-182 (INN), :185-217 (INNwProb); util.py:21-26 (build_cov_matrix);
var_state.py:25-34,36-64,66-67,76-79,94-124; sampler.py:25-34,57-63,72-86.

Known deviation (documented in DESIGN.md): flax 0.3.6 nn.Dense casts to its default
dtype=float32 inside the MLPs even with jax_enable_x64; this restatement (and the CUDA path)
keeps float64 end to end, a superset of the reference's precision.
"""
from dataclasses import dataclass, field
import math
import numpy as np
import torch

from . import threefry

torch.set_default_dtype(torch.float64)

ALPHA = 10.0  # net.py:50


@dataclass
class FlowSpec:
    dim: int
    depth: int
    hidden: tuple            # network_args["intmediate"]
    latent: str = "Gauss"    # network_args["latentSpaceName"]
    variant: str = "no_add"  # class defaults net.py:69-71: 'no_add' | 'different_add' | 'jac_eq_1' | 'add_s'
    offset: np.ndarray = None
    inds_up: list = field(default_factory=list)
    inds_down: list = field(default_factory=list)
    global_change: bool = False   # net.py:72

    def __post_init__(self):
        if self.offset is None:
            self.offset = np.zeros(self.dim)
        self.offset = np.asarray(self.offset, dtype=np.float64)

    # ---- flat parameter layout: var_state.py:106-108 over flax's sorted param dict
    def trafo_names(self):
        return ["s1", "s2", "t1", "t2"] if self.variant == "different_add" else ["s1", "s2"]

    def trafo_dims(self, name, b):
        d1, d2 = len(self.inds_up[b]), len(self.inds_down[b])
        # net.py:75-79: s1,t1 have width len(ind_down) and eat v1 (len ind_up); s2,t2 the converse
        if name in ("s1", "t1"):
            return [d1, *self.hidden, d2]
        return [d2, *self.hidden, d1]

    def layout(self):
        d = self.dim
        out = [("L", (d * (d - 1) // 2,)), ("L_diag", (d,)),
               ("dist_params", (1 if self.latent == "Student_t" else 0,)), ("mu", (d,))]
        for b in sorted(range(self.depth), key=lambda i: f"blocks_{i}"):
            if self.global_change:   # net.py:80-82; 'global_offset' < 'global_scale' < 's1' in the sorted param dict
                out.append((f"blocks_{b}/global_offset", (d,)))
                out.append((f"blocks_{b}/global_scale", (1,)))
            for tn in self.trafo_names():
                dims = self.trafo_dims(tn, b)
                for l in range(len(dims) - 1):
                    out.append((f"blocks_{b}/{tn}/Dense_{l}/bias", (dims[l + 1],)))
                    out.append((f"blocks_{b}/{tn}/Dense_{l}/kernel", (dims[l], dims[l + 1])))
        return out

    def slices(self):
        sl, start = {}, 0
        for name, shape in self.layout():
            n = int(np.prod(shape))
            sl[name] = (start, start + n, shape)
            start += n
        return sl, start

    @property
    def num_params(self):
        return self.slices()[1]


def make_index_splits(dim, depth, init_key=1, mode="jax"):
    """var_state.py:110-119: per block key,use=split(key); ind_up=choice(use, d, (d//2,), replace=False);
    ind_down=setdiff1d(arange(d), ind_up).  Returns (inds_up, inds_down, key_after)."""
    key = threefry.prng_key(init_key)
    ups, downs = [], []
    for _ in range(depth):
        key, use = threefry.split(key)
        up = threefry.choice_no_replace(use, dim, int(dim / 2))
        down = np.setdiff1d(np.arange(dim), up)
        ups.append([int(i) for i in up])
        downs.append([int(i) for i in down])
    return ups, downs, key


def init_params(spec, seed=1):
    """Reference init distributions (net.py:39-41,48-49,55-56,201-204): latent parameters and biases 0,
    hidden kernels U[-1,1), last kernel U[-1e-5,1e-5).  flax's per-module RNG folding is not
    reproducible here; draws come from numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    sl, P = spec.slices()
    theta = np.zeros(P)
    for name, (a, b, shape) in sl.items():
        if name.endswith("kernel"):
            nl = len(spec.hidden)
            last = name.split("/")[-2] == f"Dense_{nl}"
            scale = 1e-5 if last else 1.0
            theta[a:b] = 2 * scale * (rng.random(b - a) - 0.5)
        elif name.endswith("global_scale"):
            theta[a:b] = 1.0     # net.py:81
    return theta


def init_params_flax(spec, key):
    """`mynet.init(key, zeros(d))` of var_state.py:123 in flax 0.3.6 / jax 0.2.18 (both unvendored; restated from the
    published sources): module key = parent key folded with sha1(module name) (Scope.push), parameter key = module key
    folded with the creation counter (Scope.make_rng; Dense creates 'kernel' first -> 1), kernel_init = net.py:39-41 on
    jax.nn.initializers.uniform() = 0.01 * random.uniform(key, shape, float32).  `key` is the key left after the
    index-split loop (make_index_splits(...)[2]).
    PINNED: with this stream the 50 largest eigenvalues of S of the reference's stored d=6, P=411 run agree to 4e-3
    (a fresh Monte-Carlo draw moves them by 3-7 %, a different init stream by 26 %); tests/test_reference_pins.py."""
    sl, P = spec.slices()
    theta = np.zeros(P)
    nl = len(spec.hidden)
    for name, (a, b, shape) in sl.items():
        if name.endswith("global_scale"):
            theta[a:b] = 1.0     # net.py:81
        if not name.endswith("kernel"):
            continue
        parts = name.split("/")
        k = threefry.fold_in_str(key, "myINN")
        for part in parts[:-1]:
            k = threefry.fold_in_str(k, part)
        k = threefry.fold_in(k, 1)
        scale = 1e-5 if parts[-2] == f"Dense_{nl}" else 1.0
        u = threefry.uniform01(k, b - a, np.float32) * np.float32(0.01)
        theta[a:b] = (np.float32(2 * scale) * (u / np.float32(0.01) - np.float32(0.5))).astype(np.float64)
    return theta


# ---------------------------------------------------------------------------- model (torch)
def build_L(theta, spec, sl):
    """util.py:21-26 without the final L@L.T: strict upper from triu_indices(d,1) row-major + diag(exp)."""
    d = spec.dim
    a, b, _ = sl["L"]
    a2, b2, _ = sl["L_diag"]
    L = torch.zeros((d, d), dtype=theta.dtype)
    iu = torch.triu_indices(d, d, 1)
    if b > a:
        L = L.index_put((iu[0], iu[1]), theta[a:b])
    return L + torch.diag(torch.exp(theta[a2:b2]))


def latent_logpdf(y, theta, spec, sl):
    """net.py:14-20 / :27-36 with closed forms: logdet S = 2 sum L_diag,
    y^T S^-1 y = |L^-1 y|^2.  y already has the offset removed; mu subtracted here."""
    d = spec.dim
    L = build_L(theta, spec, sl)
    a, b, _ = sl["mu"]
    y = y - theta[a:b]
    w = torch.linalg.solve_triangular(L, y.unsqueeze(-1), upper=True).squeeze(-1)
    q = (w * w).sum()
    if spec.latent == "Gauss":
        a2, b2, _ = sl["L_diag"]
        return -0.5 * (d * math.log(2 * math.pi) + 2.0 * theta[a2:b2].sum() + q)
    a3, _, _ = sl["dist_params"]
    nu = torch.exp(theta[a3]) + 1.0
    # net.py:35-36 -- no -0.5 logdet S term (quirk preserved)
    return (torch.lgamma((nu + d) / 2) - torch.lgamma(nu / 2) - d / 2 * torch.log(nu * math.pi)
            - (nu + d) / 2 * torch.log(1 + q / nu))


def trafo(x, theta, spec, sl, b, tn):
    """net.py:52-61."""
    nl = len(spec.hidden)
    for l in range(nl + 1):
        a, e, shp = sl[f"blocks_{b}/{tn}/Dense_{l}/kernel"]
        a2, e2, _ = sl[f"blocks_{b}/{tn}/Dense_{l}/bias"]
        x = x @ theta[a:e].reshape(shp) + theta[a2:e2]
        x = torch.tanh(x) if l < nl else ALPHA * torch.tanh(x)
    return x


def block_forward(x, theta, spec, sl, b):
    """net.py:85-118."""
    up, down = spec.inds_up[b], spec.inds_down[b]
    u1, u2 = x[up], x[down]
    s2 = trafo(u2, theta, spec, sl, b, "s2")
    v = spec.variant
    if v == "jac_eq_1":
        v1, s2 = u1 + s2, torch.zeros_like(s2)
    elif v == "different_add":
        v1 = u1 * torch.exp(s2) + trafo(u2, theta, spec, sl, b, "t2")
    elif v == "no_add":
        v1 = u1 * torch.exp(s2)
    else:
        v1 = u1 * torch.exp(s2) + s2
    s1 = trafo(v1, theta, spec, sl, b, "s1")
    if v == "jac_eq_1":
        v2, s1 = u2 + s1, torch.zeros_like(s1)
    elif v == "different_add":
        v2 = u2 * torch.exp(s1) + trafo(v1, theta, spec, sl, b, "t1")
    elif v == "no_add":
        v2 = u2 * torch.exp(s1)
    else:
        v2 = u2 * torch.exp(s1) + s1
    out = torch.zeros_like(x)
    out = out.index_put((torch.tensor(up),), v1).index_put((torch.tensor(down),), v2)
    if spec.global_change:   # net.py:115-116
        (a, e, _), (a2, _, _) = sl[f"blocks_{b}/global_offset"], sl[f"blocks_{b}/global_scale"]
        return theta[a2] * out + theta[a:e], s2.sum() + s1.sum() + torch.log(theta[a2]) * (len(up) + len(down))
    return out, s2.sum() + s1.sum()


def block_inverse(x, theta, spec, sl, b):
    """net.py:120-153."""
    up, down = spec.inds_up[b], spec.inds_down[b]
    v1, v2 = x[up], x[down]
    s1 = trafo(v1, theta, spec, sl, b, "s1")
    v = spec.variant
    if v == "jac_eq_1":
        u2, s1 = v2 - s1, torch.zeros_like(s1)
    elif v == "different_add":
        u2 = (v2 - trafo(v1, theta, spec, sl, b, "t1")) * torch.exp(-s1)
    elif v == "no_add":
        u2 = v2 * torch.exp(-s1)
    else:
        u2 = (v2 - s1) * torch.exp(-s1)
    s2 = trafo(u2, theta, spec, sl, b, "s2")
    if v == "jac_eq_1":
        u1, s2 = v1 - s2, torch.zeros_like(s2)
    elif v == "different_add":
        u1 = (v1 - trafo(u2, theta, spec, sl, b, "t2")) * torch.exp(-s2)
    elif v == "no_add":
        u1 = v1 * torch.exp(-s2)
    else:
        u1 = (v1 - s2) * torch.exp(-s2)
    out = torch.zeros_like(x)
    out = out.index_put((torch.tensor(up),), u1).index_put((torch.tensor(down),), u2)
    if spec.global_change:   # net.py:149-150 (the affine step is undone after the inverse coupling: the reference's order)
        (a, e, _), (a2, _, _) = sl[f"blocks_{b}/global_offset"], sl[f"blocks_{b}/global_scale"]
        return (out - theta[a:e]) / theta[a2], -(s1.sum() + s2.sum()) - torch.log(theta[a2]) * (len(up) + len(down))
    return out, -(s1.sum() + s2.sum())


def inn(x, theta, spec, sl, inv=False):
    """net.py:168-182."""
    lj = torch.zeros((), dtype=x.dtype)
    order = range(spec.depth) if not inv else reversed(range(spec.depth))
    for b in order:
        x, l = (block_inverse if inv else block_forward)(x, theta, spec, sl, b)
        lj = lj + l
    return x, lj


def logp_single(x, theta, spec, sl):
    """net.py:209-213 (evaluate=True)."""
    z, lj = inn(x, theta, spec, sl, inv=False)
    return latent_logpdf(z - torch.as_tensor(spec.offset), theta, spec, sl) + lj


def sample_single(z, theta, spec, sl):
    """net.py:214-217 (evaluate=False, inv=True)."""
    p = latent_logpdf(z - torch.as_tensor(spec.offset), theta, spec, sl)
    x, lj = inn(z, theta, spec, sl, inv=True)
    return x, p - lj


class OracleState:
    """Minimal VarState restatement (var_state.py) over a flat parameter vector."""

    def __init__(self, spec, theta, sampler_seed=0):
        self.spec = spec
        self.sl, self.P = spec.slices()
        self.theta = torch.as_tensor(np.asarray(theta, dtype=np.float64)).clone()
        # sampler.py:57-60 with commSize=1, device_count=1
        k = threefry.prng_key(sampler_seed)
        k = threefry.split(k, 1)[0]
        k = threefry.split(k, 1)[0]
        self.key = k

    # var_state.py:36-43
    def logp(self, x, chunk=8192):
        x = torch.as_tensor(x)
        f = torch.func.vmap(lambda xi: logp_single(xi, self.theta, self.spec, self.sl))
        return torch.cat([f(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)])

    # var_state.py:55-64: value, coordinate gradient, flattened parameter gradient
    def eval_coordgrads(self, x, chunk=4096):
        x = torch.as_tensor(x)
        vg = torch.func.vmap(torch.func.grad_and_value(
            lambda xi, th: logp_single(xi, th, self.spec, self.sl), argnums=(0, 1)), in_dims=(0, None))
        vals, gx, gt = [], [], []
        for i in range(0, x.shape[0], chunk):
            (g0, g1), v = vg(x[i:i + chunk], self.theta)
            vals.append(v), gx.append(g0), gt.append(g1)
        return torch.cat(vals), torch.cat(gx), torch.cat(gt)

    # var_state.py:32,66-67
    def hessian(self, x, chunk=2048):
        x = torch.as_tensor(x)
        h = torch.func.vmap(torch.func.hessian(lambda xi: logp_single(xi, self.theta, self.spec, self.sl)))
        return torch.cat([h(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)])

    def cov_chol(self):
        L = build_L(self.theta, self.spec, self.sl)
        return torch.linalg.cholesky(L @ L.T)

    def latent_draw(self, n, key):
        """sampler.py:25-26,86 Gauss: mu + chol(S) xi + offset, xi = normal(key, (1,N,d)).
        Student_t (sampler.py:29-34) needs NumPy's unseeded chisquare; the oracle takes the chi^2
        stream from `self.chi2` if set (tests inject it), else raises."""
        d = self.spec.dim
        xi = torch.as_tensor(threefry.normal(key, n * d).reshape(n, d))
        a, b, _ = self.sl["mu"]
        y = xi @ self.cov_chol().T
        if self.spec.latent == "Student_t":
            a3, _, _ = self.sl["dist_params"]
            nu = float(torch.exp(self.theta[a3]) + 1.0)
            u = torch.as_tensor(self.chi2(nu, n))
            y = torch.sqrt(nu / u)[:, None] * y
        return y + self.theta[a:b] + torch.as_tensor(self.spec.offset)

    # var_state.py:76-79 + sampler.py:72-73
    def sample(self, n, chunk=8192):
        new = threefry.split(self.key, 2)
        self.key, use = new[0], new[1]
        z = self.latent_draw(n, use)
        f = torch.func.vmap(lambda zi: sample_single(zi, self.theta, self.spec, self.sl))
        xs, lps = [], []
        for i in range(0, n, chunk):
            x, lp = f(z[i:i + chunk])
            xs.append(x), lps.append(lp)
        return torch.cat(xs), torch.cat(lps), z
