"""Mirror of vmc_fluids/var_state.py: owns the ansatz and its flat parameter vector and exposes batched
evaluation, coordinate / parameter gradients, Hessian and exact sampling -- all on the sm_100a kernels.

Arrays keep the reference's (device, batch, ...) layout with a device axis of size 1 (sampler.py:26,33).
"""
import numpy as np
import torch

from . import _kernels, _capi, _threefry, net, util, global_defs, mpi_wrapper


class VarState:
    """var_state.py:18-124."""

    def __init__(self, sampler, dim, *args, network_args={}):
        self.sampler = sampler
        self.dim = dim
        self.net, self.params = self.init_net(network_args, *args)
        self._flat = net.INNwProb.flat_of(self.params)
        # var_state.py:25-27
        self.paramShapes = [(int(np.prod(shape)), tuple(shape)) for _, shape in self.net.layout()]
        self.netTreeDef = [path for path, _ in self.net.layout()]
        self.numParameters = self.net.numParameters
        self._diffusion_eq = _capi.make_equation("diffusion", {"D": 1.0})

    # ---- evaluation (var_state.py:36-67) ---------------------------------------------------------
    def _coords(self, coords):
        c = _kernels.as_dev(coords)
        if c.ndim == 2:
            c = c[None, ...]
        if c.ndim != 3 or c.shape[-1] != self.dim:
            raise ValueError("coords must have shape (devices, batch, dim)")
        return c

    def __call__(self, coords, mode="eval", avg=False):
        c = self._coords(coords)
        nd, nb = c.shape[0], c.shape[1]
        flatc = c.reshape(nd * nb, self.dim)
        if mode == "eval":  # var_state.py:38-43
            value = _kernels.logp(self.net.handle, self._flat, flatc).view(nd, nb)
            return value.mean(dim=(0, 1)) if avg else value
        if mode == "costfun":  # var_state.py:45-53: -log p and its parameter gradient (tree)
            value, _, grads = self._eval_coordgrads(flatc)
            value, grads = -value.view(nd, nb), -grads.view(nd, nb, -1)
            if avg:
                return value.mean(dim=(0, 1)), self.net.tree_from_flat(grads.mean(dim=(0, 1)).contiguous())
            return value, grads
        if mode == "eval_coordgrads":  # var_state.py:55-64
            value, cg, pg = self._eval_coordgrads(flatc)
            return value.view(nd, nb), cg.view(nd, nb, self.dim), pg.view(nd, nb, -1)
        raise ValueError(f"unknown mode {mode!r}")

    def _eval_coordgrads(self, flatc):
        h = self.net.handle
        n = flatc.shape[0]
        O = _kernels.empty(max(n, 1), h.Pp)
        out = _kernels.local_terms(h, self._flat, flatc, self._diffusion_eq, O=O, ldo=h.Pp, want=("logp", "grad"))
        return out["logp"], out["grad"], O[:n, :h.P]

    def hessian(self, coords):
        """var_state.py:66-67 -> (devices, batch, dim, dim)."""
        c = self._coords(coords)
        H = _kernels.hessian(self.net.handle, self._flat, c.reshape(-1, self.dim))
        return H.view(c.shape[0], c.shape[1], self.dim, self.dim)

    def real_space_prob(self, x, params):
        """var_state.py:73-74."""
        return self.net.apply(params, x)

    # ---- sampling (var_state.py:76-79) -----------------------------------------------------------
    def latent_dist_params(self):
        p = self.params["params"]
        return {"S": util.build_cov_matrix(p["L"], p["L_diag"], self.dim), "mu": p["mu"], "dist_params": p["dist_params"]}

    def chi2_draws(self, n, first=0, n_total=None):
        """sampler.py:32: chi^2(nu) variates from NumPy's global RNG (host), Student-t latent only.  Every rank draws the
        whole numSamples vector (identically seeded ranks then hold identical vectors) and keeps the entries of its own
        global sample indices [first, first + n), so the sample set does not depend on the number of ranks."""
        if self.net.latentSpaceName != "Student_t":
            return None
        nu = float(torch.exp(self.params["params"]["dist_params"][0]) + 1.0)
        n_total = n if n_total is None else n_total
        return _kernels.as_dev(np.random.chisquare(nu, size=(n_total,))[first:first + n])

    def sample_range(self, key, first, n, n_total, chi2=None):
        """Samples [first, first+n) of the n_total-sample stream of `key`: (x (n,d), logp (n,))."""
        return _kernels.sample(self.net.handle, self._flat, key, first, n, n_total, chi2)

    def sample(self, numSamples):
        """var_state.py:76-79 -> (coords (1, n_local, d), logp (1, n_local)); with several ranks each draws its own
        contiguous slice of the single global stream (see sampler.py docstring)."""
        key = self.sampler.next_key()
        first, n = mpi_wrapper.shard_range(numSamples)
        x, lp = self.sample_range(key, first, n, numSamples, self.chi2_draws(n, first, numSamples))
        return x[None, ...], lp[None, ...]

    def average_tree(self, tree, axis=(0, 1)):
        """var_state.py:81-86."""
        return {k: (self.average_tree(v, axis) if isinstance(v, dict) else v.mean(dim=axis)) for k, v in tree.items()}

    def integrate(self, grid):
        """var_state.py:88-91."""
        coords = _kernels.as_dev(grid.coords)[None, ...]
        return float((grid.bin_area * torch.exp(self(coords))).sum())

    # ---- parameter utilities (var_state.py:94-108) -----------------------------------------------
    def set_parameters(self, p_new):
        p = _kernels.as_dev(p_new).reshape(-1)
        if p.numel() != self.numParameters:
            raise ValueError(f"expected {self.numParameters} parameters, got {p.numel()}")
        self._flat = p.clone()
        self.params = self.net.tree_from_flat(self._flat)

    def get_parameters(self):
        return self.flatten_tree(self.params)

    def flatten_tree(self, tree):
        leaves = []

        def walk(node):
            for k in sorted(node.keys()):
                v = node[k]
                walk(v) if isinstance(v, dict) else leaves.append(v.reshape(-1).to(torch.float64))
        walk(tree)
        return torch.cat(leaves).clone() if leaves else _kernels.zeros(0)

    def init_net(self, network_args, key, depth, **kwargs):
        """var_state.py:110-124: per block key,use = split(key); ind_up = choice(use, d, (d//2,), replace=False);
        ind_down = setdiff1d(arange(d), ind_up); then params = net.init(key, zeros(d))."""
        key = _threefry.PRNGKey(key)
        inds_up, inds_down = [], []
        for _ in range(depth):
            key, use_key = _threefry.split(key)
            ind_up = _threefry.choice_no_replace(use_key, self.dim, int(self.dim / 2))
            ind_down = np.setdiff1d(np.arange(self.dim), ind_up)
            inds_up.append(ind_up)
            inds_down.append(ind_down)
        mynet = net.INNwProb(inds_up, inds_down, **network_args)
        params = mynet.init(key, np.zeros(self.dim))
        return mynet, params
