"""Mirror of vmc_fluids/sampler.py: latent-space draws for the flow ansatz.

Only the exact generators are live in the reference (the Metropolis branch calls a field that
does not exist, sampler.py:95), so only they are built.  Key chain as sampler.py:57-60,73; the bulk normal
draws run on the device in JAX's threefry2x32 counter layout.

Deliberate difference for commSize > 1: the reference derives a per-rank key and lets every rank draw the
full numSamples (over-counting); here all ranks share ONE global stream and draw disjoint index
ranges, so results do not depend on the number of GPUs.
"""
from dataclasses import dataclass
import numpy as np
import torch

from . import _kernels, _threefry, global_defs, mpi_wrapper


def radial_update_prop(key, numChains, mcmc_info):
    raise NotImplementedError("Metropolis sampling is dead code in the reference (sampler.py:88-111) and not built")


def gauss_sample(key, dist_params, numSamples):
    """sampler.py:25-26: multivariate_normal(key, mu, S, (1, N)) = mu + chol(S) xi."""
    S, mu = dist_params["S"], dist_params["mu"]
    d = mu.shape[0]
    xi = _kernels.normal(key, 0, numSamples * d, numSamples * d).view(numSamples, d)
    return (mu + xi @ torch.linalg.cholesky(S).T)[None, ...]


def student_t_sample(key, dist_params, numSamples):
    """sampler.py:29-34 (chi^2 from NumPy's global RNG on the host, exactly as the reference)."""
    nu = float(torch.exp(dist_params["dist_params"][0]) + 1e0)
    u = _kernels.as_dev(np.random.chisquare(nu, size=(numSamples,)))
    S, mu = dist_params["S"], dist_params["mu"]
    d = mu.shape[0]
    xi = _kernels.normal(key, 0, numSamples * d, numSamples * d).view(numSamples, d)
    y = xi @ torch.linalg.cholesky(S).T
    return (torch.sqrt(nu / u)[:, None] * y + mu)[None, ...]


@dataclass
class Sampler:
    """sampler.py:48-86."""
    key: int = 0
    numChains: int = 1
    dim: int = 2
    name: str = "Gauss"
    updateProposer: callable = radial_update_prop
    mcmc_info: any = None

    def __post_init__(self):
        k = _threefry.PRNGKey(self.key)
        # sampler.py:58-59 with ONE global stream (rank 0's key on every rank, see module docstring)
        k = _threefry.split(k, 1)[0]
        k = _threefry.split(k, global_defs.device_count())[0]
        self.key = k
        self.exact_samples = self.name in ["Gauss", "Student_t"]
        self.exact_sample_generator_dict = {"Gauss": gauss_sample, "Student_t": student_t_sample}
        self.states = None
        if self.mcmc_info is None:
            self.mcmc_info = {"offset": np.zeros(self.dim)}
        if not self.exact_samples:
            raise NotImplementedError(f"latent distribution {self.name!r}: only the exact samplers 'Gauss' and "
                                      "'Student_t' run in the reference (sampler.py:61,95)")

    def next_key(self):
        """sampler.py:73: self.key, key_to_use = split(self.key, 2)."""
        ks = _threefry.split(self.key, 2)
        self.key = ks[0]
        return ks[1]

    def offset_tensor(self):
        return _kernels.as_dev(self.mcmc_info["offset"])

    def __call__(self, numSamples, dist_params, multipleOf=1):
        """sampler.py:72-86: latent samples of shape (1, numSamples, dim)."""
        key_to_use = self.next_key()
        z = self.exact_sample_generator_dict[self.name](key_to_use, dist_params, numSamples)
        return z + self.offset_tensor()[None, None, :]
