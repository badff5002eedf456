"""Mirror of vmc_fluids/grid.py: the regular evaluation grid main.py builds for d = 2 (main.py:100-104) and hands
to VarState.integrate (var_state.py:88-91) and the plotting helpers.

Same constructor and attributes (sym, dim, bounds, n_gridpoints, widths, bin_area, range, vals, meshgrid, coords);
host-side NumPy geometry only -- the density on the grid points is evaluated by the logp kernel."""
import numpy as np


class Grid:
    """grid.py:7-29.  `bounds[k]` is the half-width of axis k when `sym` (cells span [-b, b)), else its length
    (cells span [0, b)); every axis gets `n_gridpoints` cells; `coords` lists the cell origins, axis 1 fastest over
    axis 0 in the order of np.meshgrid's default 'xy' indexing, exactly as the reference flattens them."""

    def __init__(self, bounds, n_gridpoints, sym=True):
        bounds = np.asarray(bounds, dtype=np.float64)
        self.sym = sym
        self.dim = bounds.shape[0]
        self.bounds = bounds
        self.n_gridpoints = n_gridpoints
        self.widths = (2.0 if sym else 1.0) * bounds / n_gridpoints
        self.bin_area = np.prod(self.widths)
        lo = -bounds if sym else np.zeros_like(bounds)
        self.range = [[float(a), float(b)] for a, b in zip(lo, bounds)]
        self.vals = [np.arange(a, b, w) for a, b, w in zip(lo, bounds, self.widths)]
        self.meshgrid = np.meshgrid(*self.vals)
        self.coords = np.stack(self.meshgrid, axis=-1).reshape(n_gridpoints ** self.dim, self.dim)
