"""Mirror of vmc_fluids/util.py: build_cov_matrix (util.py:21-26), Timings (:35-52), store_infos (:29-32)."""
import time
import numpy as np
import torch


def build_cov_matrix(L_para, L_diag, dim):
    """util.py:21-26: strict upper triangle (triu_indices(dim,1), row-major) + diag(exp(L_diag)); S = L L^T."""
    L = torch.zeros((dim, dim), dtype=torch.float64, device=L_diag.device)
    iu = torch.triu_indices(dim, dim, 1, device=L_diag.device)
    if iu.numel():
        L[iu[0], iu[1]] = L_para.to(torch.float64)
    L = L + torch.diag(torch.exp(L_diag.to(torch.float64)))
    return L @ L.T


def store_infos(wdir, infos, name="infos.hdf5"):
    """util.py:29-32 writes one HDF5 dataset per key.  h5py is used when importable; otherwise the same keys go
    into `<name>.npz` (this image has no h5py)."""
    data = {k: np.asarray([np.asarray(torch.as_tensor(v).cpu()) for v in vals]) for k, vals in infos.items()}
    try:
        import h5py
        with h5py.File(wdir + name, "w") as f:
            for key, value in data.items():
                f.create_dataset(key, data=value)
    except ImportError:
        np.savez(wdir + name + ".npz", **data)


class Timings():
    """util.py:35-52."""

    def __init__(self):
        self.timing_dict = {}

    def start_timing(self, key):
        if key not in self.timing_dict.keys():
            self.timing_dict[key] = []
        self.timing_dict[key].append(- time.perf_counter())

    def stop_timing(self, key):
        self.timing_dict[key][-1] += time.perf_counter()

    def print_timings(self):
        total = 0
        for key, value in self.timing_dict.items():
            print(f"\t > {key}: {value[-1]}")
            total += value[-1]
        print(f"\t > TOTAL: {total}")
