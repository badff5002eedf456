"""Mirror of vmc_fluids/net.py: the invertible-network ansatz with a latent Gauss / Student-t density.

The reference builds flax modules and lets JAX trace them; here the classes only DESCRIBE the architecture
(index splits, widths, coupling variant, latent density) and `apply` dispatches to the sm_100a kernels.
Parameters are a nested dict with flax's names, {'params': {'L', 'L_diag', 'dist_params', 'mu',
'myINN': {'blocks_i': {['global_offset', 'global_scale',] 's1'|'s2'|'t1'|'t2': {'Dense_k': {'bias', 'kernel'}}}}}}, whose leaves are VIEWS into one
flat float64 device vector in the reference's flatten order (var_state.py:106-108).
"""
from dataclasses import dataclass, field
import numpy as np
import torch

from . import _kernels, _threefry

ALPHA = 1e1  # net.py:50


class SingleBlock:
    """Coupling-variant switches.  The reference selects the variant by editing these class defaults
    (net.py:69-72; main.py:43-58); INNwProb reads them at construction unless `variant=` is given."""
    jac_eq_1: bool = False
    different_add: bool = False
    no_add: bool = True
    global_change: bool = False  # net.py:72,80-82: per-block global_scale / global_offset (see INNwProb.global_change)

    @classmethod
    def variant_name(cls):
        if cls.jac_eq_1:
            return "jac_eq_1"
        if cls.different_add:
            return "different_add"
        if cls.no_add:
            return "no_add"
        return "add_s"


def uniform_init(u01, scale=1.0):
    """net.py:39-41 applied to U[0,1) draws: 2*scale*(u - 0.5)."""
    return 2.0 * scale * (u01 - 0.5)


@dataclass
class INNwProb:
    """net.py:185-217."""
    inds_up: list
    inds_down: list
    intmediate: tuple = (3,)
    offset: any = None
    latentSpaceName: str = "Gauss"
    dim: int = 2
    variant: str = None
    global_change: bool = None     # None: SingleBlock.global_change.  The reference's inverse branch (net.py:149-150) undoes
                                   # the affine step after the inverse coupling, so with scale != 1 or offset != 0 sampling no
                                   # longer draws from the evaluated density; kept exactly as the reference has it.

    def __post_init__(self):
        self.intmediate = tuple(int(h) for h in self.intmediate)
        self.offset = np.zeros(self.dim) if self.offset is None else np.asarray(
            self.offset.detach().cpu() if isinstance(self.offset, torch.Tensor) else self.offset, dtype=np.float64)
        if self.latentSpaceName not in ("Gauss", "Student_t"):
            raise KeyError(self.latentSpaceName)  # net.py:197-198 has only these two
        if self.variant is None:
            self.variant = SingleBlock.variant_name()
        if self.global_change is None:
            self.global_change = bool(SingleBlock.global_change)
        self.depth = len(self.inds_up)
        self.inds_up = [[int(i) for i in np.asarray(u).ravel()] for u in self.inds_up]
        self.inds_down = [[int(i) for i in np.asarray(u).ravel()] for u in self.inds_down]
        self.handle = _kernels.FlowHandle(self.dim, self.depth, self.intmediate, self.variant, self.latentSpaceName,
                                          self.inds_up, self.inds_down, self.offset, self.global_change)
        self.numParameters = self.handle.P

    # ---- flat layout ---------------------------------------------------------------------------
    def trafo_names(self):
        return ["s1", "s2", "t1", "t2"] if self.variant == "different_add" else ["s1", "s2"]

    def layout(self):
        """[(path tuple, shape)] in flatten order: sorted keys, ravel()ed leaves."""
        d = self.dim
        d1, d2 = d // 2, d - d // 2
        out = [(("params", "L"), (d * (d - 1) // 2,)), (("params", "L_diag"), (d,)),
               (("params", "dist_params"), (1 if self.latentSpaceName == "Student_t" else 0,)), (("params", "mu"), (d,))]
        for b in sorted(range(self.depth), key=lambda i: f"blocks_{i}"):
            if self.global_change:      # 'global_offset' < 'global_scale' < 's1'
                out.append((("params", "myINN", f"blocks_{b}", "global_offset"), (d,)))
                out.append((("params", "myINN", f"blocks_{b}", "global_scale"), (1,)))
            for tn in self.trafo_names():
                dims = [d1, *self.intmediate, d2] if tn in ("s1", "t1") else [d2, *self.intmediate, d1]
                for l in range(len(dims) - 1):
                    out.append((("params", "myINN", f"blocks_{b}", tn, f"Dense_{l}", "bias"), (dims[l + 1],)))
                    out.append((("params", "myINN", f"blocks_{b}", tn, f"Dense_{l}", "kernel"), (dims[l], dims[l + 1])))
        return out

    def tree_from_flat(self, flat):
        """Nested dict whose leaves are views of `flat`."""
        tree, start = {}, 0
        for path, shape in self.layout():
            n = int(np.prod(shape))
            node = tree
            for k in path[:-1]:
                node = node.setdefault(k, {})
            node[path[-1]] = flat[start:start + n].view(shape)
            start += n
        assert start == self.numParameters
        return tree

    def init(self, key, x=None):
        """flax 0.3.6 `Module.init(key, x)`: latent parameters and biases zero, hidden kernels U[-1,1), last kernel
        U[-1e-5,1e-5) (net.py:39-41,48-49,55-56,201-204), drawn from flax's RNG tree: the key of a module is its
        parent's key folded with the SHA-1 of the module name (`Scope.push`), the key of the k-th parameter created in
        a module is the module key folded with k (`Scope.make_rng`; a Dense layer creates its kernel first), and
        jax.nn.initializers.uniform draws float32.  With this stream the spectrum of S of the reference's stored d=6
        run is reproduced (tests/test_reference_pins.py)."""
        key = _threefry.PRNGKey(key) if np.isscalar(key) else np.asarray(key, dtype=np.uint32)
        flat = np.zeros(self.numParameters)
        start = 0
        nl = len(self.intmediate)
        for path, shape in self.layout():
            n = int(np.prod(shape))
            if path[-1] == "kernel":
                k = key
                for name in path[1:-1]:                       # 'myINN', 'blocks_b', 's1', 'Dense_l'
                    k = _threefry.fold_in_str(k, name)
                k = _threefry.fold_in(k, 1)
                scale = 1e-5 if path[-2] == f"Dense_{nl}" else 1.0
                u = _threefry.uniform01_f32(k, n) * np.float32(0.01)      # initializers.uniform(): scale 1e-2, float32
                flat[start:start + n] = (np.float32(2.0 * scale) * (u / np.float32(0.01) - np.float32(0.5))).astype(np.float64)
            elif path[-1] == "global_scale":
                flat[start] = 1.0                                         # jax.nn.initializers.ones (net.py:81)
            start += n
        flat_t = _kernels.as_dev(flat)
        return self.tree_from_flat(flat_t)

    # ---- evaluation -----------------------------------------------------------------------------
    @staticmethod
    def flat_of(params):
        """The flat vector behind a parameter tree built by tree_from_flat (or a re-flattening of any tree)."""
        leaves = []

        def walk(node):
            for k in sorted(node.keys()):
                v = node[k]
                walk(v) if isinstance(v, dict) else leaves.append(v.reshape(-1))
        walk(params)
        base = leaves[0]._base if leaves and leaves[0]._base is not None else None
        if base is not None and all(l._base is base for l in leaves if l.numel()) and base.is_contiguous():
            return base
        return torch.cat([l.to(torch.float64) for l in leaves]) if leaves else _kernels.zeros(0)

    def apply(self, params, x, evaluate=True, inv=False):
        """net.py:209-217 for one point (d,) or a batch (n, d)."""
        theta = self.flat_of(params)
        xt = _kernels.as_dev(x)
        single = xt.ndim == 1
        xb = xt.reshape(-1, self.dim)
        if evaluate:
            if inv:
                y, lj, _ = _kernels.transform(self.handle, theta, xb, True)
                yy, _, lat = _kernels.transform(self.handle, theta, y, False, want_latent=True)  # latent pdf of y
                out = lat + lj
            else:
                out = _kernels.logp(self.handle, theta, xb)
            return out[0] if single else out
        y, lj, lat = _kernels.transform(self.handle, theta, xb, bool(inv), want_latent=True)
        val = lat - lj
        return (y[0], val[0]) if single else (y, val)


# the reference also defines SanityINN (net.py:220-235), a one-parameter rescaling with INN's (x, log_jac) interface; its only
# use is a commented-out line (var_state.py:122) where it could not replace INNwProb (VarState needs a log-probability), so
# it is not built here.
