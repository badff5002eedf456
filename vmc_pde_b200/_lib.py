"""Loader of libvmcpde.so (the C-ABI of include/vmcpde.h) via ctypes.

There is no CPU fallback: if the shared library is missing or CUDA is unavailable the calls raise.
PyTorch is used only for device memory, streams and torch.distributed plumbing.
"""
import ctypes as C
import os

from ._capi import FlowConfig, Equation

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvmcpde.so")
_lib = None

_vp, _i64, _i32, _u32, _dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_double

# name -> (restype, argtypes): mirrors include/vmcpde.h one to one
SIGNATURES = {
    "vmcpde_last_error": (C.c_char_p, []),
    "vmcpde_version": (C.c_int, []),
    "vmcpde_padded_params": (_i32, [_i32]),
    "vmcpde_flow_create": (C.c_int, [C.POINTER(FlowConfig), C.POINTER(_vp)]),
    "vmcpde_flow_destroy": (None, [_vp]),
    "vmcpde_flow_num_params": (_i32, [_vp]),
    "vmcpde_flow_param_offsets": (C.c_int, [_vp, C.POINTER(_i32)]),
    "vmcpde_sample": (C.c_int, [_vp, _vp, _u32, _u32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "vmcpde_normal": (C.c_int, [_u32, _u32, _i64, _i64, _i64, _vp, _vp]),
    "vmcpde_uniform": (C.c_int, [_u32, _u32, _i64, _i64, _i64, _vp, _vp]),
    "vmcpde_logp": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "vmcpde_local_terms": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(Equation), _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "vmcpde_flow_transform": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "vmcpde_hessian": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "vmcpde_moments_workspace_bytes": (C.c_int, [_i64, _i64, C.POINTER(C.c_size_t)]),
    "vmcpde_moments1": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_center_force": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_gram_matvec": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_gram": (C.c_int, [_vp, _i64, _i64, _i32, _i32, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "vmcpde_gram_split_workspace_bytes": (C.c_int, [_i64, _i32, C.POINTER(C.c_size_t)]),
    "vmcpde_gram_split": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_packed_tiles_len": (_i64, [_i32]),
    "vmcpde_pack_upper_tiles": (C.c_int, [_vp, _i32, _vp, _vp]),
    "vmcpde_unpack_upper_tiles": (C.c_int, [_vp, _i32, _vp, _vp]),
    "vmcpde_allreduce_sum": (C.c_int, [_vp, _vp, _i64, _vp]),
    "vmcpde_allreduce_moments": (C.c_int, [_vp, C.POINTER(_vp), _i32, _i32, _vp, _i64, _vp, _vp]),
    "vmcpde_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "vmcpde_nccl_comm_init": (C.c_int, [_i32, _i32, C.c_char_p, C.POINTER(_vp)]),
    "vmcpde_nccl_comm_destroy": (C.c_int, [_vp]),
    "vmcpde_sym_finalize": (C.c_int, [_vp, _i32, _dbl, _vp]),
    "vmcpde_diag_shift": (C.c_int, [_vp, _vp, _i32, _i32, _dbl, _vp]),
    "vmcpde_dmma_probe": (C.c_int, [_vp, _i32, C.POINTER(_dbl), _vp]),
    "vmcpde_gemm_tn_splitk": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i64, _i32, _vp]),
    "vmcpde_syrk_tn": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i64, _dbl, _dbl, _vp]),
    "vmcpde_gemm_tn": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i64, _dbl, _dbl, _vp]),
    "vmcpde_eigh_workspace_bytes": (C.c_int, [_i32, _i32, C.POINTER(C.c_size_t)]),
    "vmcpde_eigh": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_eigh_cols": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, C.c_size_t, _vp]),
    "vmcpde_eigh_factor": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_eigh_backtransform": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, C.c_size_t, _vp]),
    "vmcpde_solve_tail_range": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _dbl, _dbl, _dbl, _i32, _i32, _i32,
                                          _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_eigh_launch_count": (C.c_int, [_i32, _i32, C.POINTER(_i32)]),
    "vmcpde_solve_tail_workspace_bytes": (C.c_int, [_i32, _i32, C.POINTER(C.c_size_t)]),
    "vmcpde_solve_tail": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _dbl, _dbl, _dbl, _i32, _dbl,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vmcpde_chol_solve": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "vmcpde_solve_scalars": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _dbl, _vp, _vp, _vp]),
    "vmcpde_observables_workspace_bytes": (C.c_int, [_i32, C.POINTER(C.c_size_t)]),
    "vmcpde_obs_first": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "vmcpde_obs_central": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "vmcpde_ball_points": (C.c_int, [_u32, _u32, _i64, _i64, _i64, _i32, _dbl, _vp, _vp]),
    "vmcpde_particles_step": (C.c_int, [_vp, _i64, _i32, _dbl, _i32, _i32, C.POINTER(Equation), _u32, _u32, _vp]),
    "vmcpde_sum_exp": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
}
# filled in as later translation units land (solve / eigh / observables)
OPTIONAL_SIGNATURES = {}


def load():
    """Return the loaded library, raising loudly when it cannot be used."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m vmc_pde_b200.build`. "
                           "vmc_pde_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in {**SIGNATURES, **OPTIONAL_SIGNATURES}.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().vmcpde_last_error()
        raise RuntimeError(f"libvmcpde error {rc}: {msg.decode() if msg else ''}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("vmc_pde_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    from . import global_defs
    global_defs.device()


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    """torch's current stream ON THIS PROCESS'S device (global_defs.device() also makes that device current)."""
    import torch
    from . import global_defs
    return C.c_void_p(torch.cuda.current_stream(global_defs.device()).cuda_stream)


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
