"""ctypes view of include/vmcpde.h (structures shared by the product loader and the test tooling)."""
import ctypes as C
import numpy as np

VARIANTS = {"no_add": 0, "different_add": 1, "jac_eq_1": 2, "add_s": 3}
GLOBAL_CHANGE = 0x100   # VMCPDE_GLOBAL_CHANGE, OR-ed into the variant
LATENTS = {"Gauss": 0, "Student_t": 1}
EQUATIONS = {"diffusion": 0, "diffusion_drift": 1, "diffusion_anisotropic": 2,
             "advection_hamiltonian": 3, "advection_paper": 4, "advection_hamiltonian_wDiss": 5}


class FlowConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("depth", C.c_int32), ("n_hidden_layers", C.c_int32),
                ("hidden", C.c_int32), ("variant", C.c_int32), ("latent", C.c_int32),
                ("ind_up", C.POINTER(C.c_int32)), ("ind_down", C.POINTER(C.c_int32)),
                ("offset", C.POINTER(C.c_double)), ("hidden_widths", C.POINTER(C.c_int32))]


class Equation(C.Structure):
    _fields_ = [("mode", C.c_int32), ("D", C.c_double), ("mu", C.c_double), ("m", C.c_double),
                ("omega", C.c_double), ("lam", C.c_double), ("T", C.c_double), ("gamma", C.c_double),
                ("t", C.c_double), ("tangents", C.c_void_p)]


def make_flow_config(dim, depth, hidden, variant, latent, inds_up, inds_down, offset, global_change=False):
    """Returns (FlowConfig, keepalive) -- keepalive owns the host arrays the struct points to."""
    hidden = tuple(hidden)
    up = np.ascontiguousarray(np.asarray(inds_up, dtype=np.int32).reshape(-1))
    down = np.ascontiguousarray(np.asarray(inds_down, dtype=np.int32).reshape(-1))
    off = np.ascontiguousarray(np.asarray(offset, dtype=np.float64).reshape(-1))
    if off.size != dim:
        raise ValueError("offset must have `dim` entries")
    hw = np.ascontiguousarray(np.asarray(hidden if hidden else (0,), dtype=np.int32))
    cfg = FlowConfig(dim, depth, len(hidden), hidden[0] if hidden else 0,
                     (VARIANTS[variant] if isinstance(variant, str) else int(variant)) | (GLOBAL_CHANGE if global_change else 0),
                     LATENTS[latent] if isinstance(latent, str) else int(latent),
                     up.ctypes.data_as(C.POINTER(C.c_int32)), down.ctypes.data_as(C.POINTER(C.c_int32)),
                     off.ctypes.data_as(C.POINTER(C.c_double)), hw.ctypes.data_as(C.POINTER(C.c_int32)))
    return cfg, (up, down, off, hw)


def make_equation(name, params, t=0.0, tangents_ptr=None):
    p = dict(D=1.0, mu=0.0, m=1.0, omega=1.0, lam=0.0, T=1.0, gamma=0.0)
    p.update({k: float(v) for k, v in params.items() if k in p})
    return Equation(EQUATIONS[name], p["D"], p["mu"], p["m"], p["omega"], p["lam"], p["T"], p["gamma"],
                    float(t), tangents_ptr)
