"""util.py of the reference, re-hosted: build_cov_matrix (util.py:21-26), store_infos (:29-32) and Timings (:35-52; same
interface, device-synchronised sections)."""
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
    """util.py:29-32: one HDF5 dataset per key of `infos` (lists of per-step values) in `wdir + name`.
    h5py is used when importable; otherwise the built-in writer (_hdf5.py) produces the same flat-group file to the HDF5
    specification for the reference's plotting scripts (visualization.py:141-280, paper_plot/*.py); a read-back by h5py
    itself could not be tried in the build image (no libhdf5), the built-in reader handles both."""
    def as_np(v):
        if isinstance(v, torch.Tensor):
            return v.detach().cpu().numpy()
        return np.asarray(v)
    data = {}
    for k, vals in infos.items():
        data[k] = np.asarray([as_np(v) for v in vals]) if isinstance(vals, (list, tuple)) else as_np(vals)
        if data[k].dtype == object:       # e.g. snr is None for the Cholesky solver: store what h5py could not either
            raise TypeError(f"infos[{k!r}] is ragged or holds None entries")
    try:
        import h5py
    except ImportError:
        from . import _hdf5
        _hdf5.write(wdir + name, data)
        return
    with h5py.File(wdir + name, "w") as f:
        for key, value in data.items():
            f.create_dataset(key, data=value)


def load_infos(wdir, name="infos.hdf5"):
    """Reads a file written by store_infos (or by the reference) back as {key: ndarray}; also the way to RESUME a run:
    `infos["parameters"][-1]` holds the flat parameter vector when the driver logs it (the reference stores none)."""
    try:
        import h5py
    except ImportError:
        from . import _hdf5
        return _hdf5.read(wdir + name)
    with h5py.File(wdir + name, "r") as f:
        return {k: np.array(f[k]) for k in f.keys()}


def save_checkpoint(path, vState, t, stepper=None, infos=None):
    """Everything a run needs to continue bit for bit (the reference has no checkpoints): the flat
    parameter vector, the sampler key (the only RNG state of the exact samplers), the time and the stepper's dt, written as
    one flat-group HDF5 file next to infos.hdf5.  `infos` (the driver's history dict) is stored under "infos.<key>" names."""
    data = {"parameters": vState.get_parameters().detach().cpu().numpy(), "sampler_key": np.asarray(vState.sampler.key, dtype=np.uint32),
            "time": np.float64(t)}
    if stepper is not None:
        data["stepper_dt"] = np.float64(stepper.dt)
    if infos:
        for k, vals in infos.items():
            data["infos." + k] = np.asarray([v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v) for v in vals])
    from . import _hdf5
    _hdf5.write(path, data)


def load_checkpoint(path, vState, stepper=None):
    """Restores what save_checkpoint wrote into `vState` (parameters, sampler key) and `stepper` (dt); returns (t, infos)."""
    from . import _hdf5
    d = _hdf5.read(path)
    vState.set_parameters(torch.as_tensor(d["parameters"]))
    vState.sampler.key = d["sampler_key"].astype(np.uint32)
    if stepper is not None and "stepper_dt" in d:
        stepper.dt = float(d["stepper_dt"])
    infos = {k[len("infos."):]: list(v) for k, v in d.items() if k.startswith("infos.")}
    return float(d["time"]), infos


class Timings:
    """Wall-clock sections by name (the interface of util.py:35-52: `timing_dict[name]` is the list of durations in seconds,
    one per start/stop pair; `print_timings` reports the latest of each and their sum in the reference's format).
    Kernels are launched asynchronously, so a section is closed only after the device has drained."""

    def __init__(self):
        self.timing_dict = {}
        self._open = {}

    @staticmethod
    def _now():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return time.perf_counter()

    def start_timing(self, key):
        self.timing_dict.setdefault(key, []).append(0.0)
        self._open[key] = self._now()

    def stop_timing(self, key):
        self.timing_dict[key][-1] = self._now() - self._open.pop(key)

    def print_timings(self):
        latest = {name: runs[-1] for name, runs in self.timing_dict.items()}
        for name, seconds in latest.items():
            print(f"\t > {name}: {seconds}")
        print(f"\t > TOTAL: {sum(latest.values())}")
