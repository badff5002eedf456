"""vmc_pde_b200 -- B200 (sm_100a) implementation of vmc_pde's TDVP time-step hot path.

Module names mirror the reference's flat layout (vmc_fluids/*.py): sampler, net, var_state, evolutionEq, tdvp,
stepper, mpi_wrapper, util, global_defs.  Everything numerical runs in libvmcpde.so (include/vmcpde.h); there is no
CPU fallback -- importing is cheap, but any computation raises if the library or a CUDA device is missing.
"""
__all__ = ["sampler", "net", "var_state", "evolutionEq", "tdvp", "stepper", "mpi_wrapper", "util", "global_defs"]
