"""Mirror of vmc_fluids/global_defs.py: dtype aliases and device selection.

The reference pmaps over jax.devices(); here one process owns one B200 (torch.distributed rank ->
cuda:LOCAL_RANK) and there is no pmap axis, so arrays carry a leading device axis of size 1 exactly as
the reference's exact samplers do (sampler.py:26,33).
"""
import os
import numpy as np
import torch

tCpx = np.complex128   # global_defs.py:6
tReal = np.float64     # global_defs.py:8
dtype = torch.float64


def _local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


_bound = None


def device():
    """The CUDA device of this process (global_defs.py:16-20 picks devices[rank % ndev]).  The first call also makes it
    the CURRENT device of the process: the C-ABI launches kernels on the current device / torch's current stream, so a
    multi-GPU user does not need a torch.cuda.set_device of their own."""
    global _bound
    if _bound is None:
        if not torch.cuda.is_available():
            raise RuntimeError("vmc_pde_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
        _bound = torch.device("cuda", _local_rank() % torch.cuda.device_count())
    if torch.cuda.current_device() != _bound.index:
        torch.cuda.set_device(_bound)
    return _bound


myPmapDevices = None


def devices():
    return [device()]


def device_count():
    """global_defs.py:47-48; always 1 (one process per GPU)."""
    return 1


def set_pmap_devices(devs):
    """global_defs.py:36-44 kept for source compatibility; a process drives exactly one GPU."""
    devs = list(devs) if isinstance(devs, (list, tuple)) else [devs]
    if len(devs) != 1:
        raise ValueError("vmc_pde_b200 drives one GPU per process; launch one rank per GPU instead")


def pmap_for_my_devices(fun, *args, **kwargs):
    """global_defs.py:24; there is no tracing compiler here -- returns `fun` applied over the size-1 device axis."""
    def wrapped(*a):
        return fun(*[x[0] if hasattr(x, "ndim") and x.ndim > 0 else x for x in a])[None, ...]
    return wrapped
