"""Extracts the reference-held results -- the five stored runs under /root/reference/vmc_fluids/paper_plot/data_* --
into small fixtures tests/golden/ref_*.npz.  Run from the repository root IN THIS CONTAINER (the GPU box has no
/root/reference):

    python tests/golden/make_reference_pins.py

These are the only numbers the reference repository itself holds for the hot path: outputs of its own
JAX runs, written by util.store_infos (util.py:29-32).  The HDF5 files are parsed by vmc_pde_b200/_hdf5.py (h5py is
not installed).  What each file pins (tests/test_reference_pins.py, tests/test_gpu_reference_pins.py):

  ref_wiener_T10.npz   exact_dyn.py __main__ ("hamiltonian" case, N=10^4, dt=1e-2, 1201 records): every record is a
                       deterministic function of jax.random.{PRNGKey,split,normal} and the integrator, so the CPU
                       restatement and the GPU particle kernel must reproduce x1, covar and the ball counts to round-off;
  ref_wiener_Tdiff.npz same script with edited parameters (coupled oscillators, unequal temperatures -- not in the
                       checked-in code); only its t=0 record (the initial normal draw) is usable;
  ref_inn_Tdiff.npz    main.py mode harmonicOsc_diff (d=6, P=411, different_add, Heun, dt=1e-4*1.3); same edited
                       physics, so only the first record is usable: it is the second right-hand side of the first Heun
                       step, i.e. it pins the sampler key chain, the multivariate_normal layout and the ball-integral
                       draws (tdvp.py:143-162) far below the Monte-Carlo scatter;
  ref_diff8_gauss.npz  main.py mode diffusion with a Gauss latent (d=8; the stored run has P=392, the checked-in
                       architecture 364): first record pins the first sampler draw; the trajectory is free diffusion,
                       covar(t) = covar(0) + 2t, entropy(t) = 4 log(2 pi e (1+2t)) (visualization.py:188);
  ref_diff8_student.npz same with the Student_t latent (host chi^2 is unseeded: bands only).
  ref_wiener_Tdiff_infos.hdf5  the stored file of ref_wiener_Tdiff itself, byte for byte (written by the reference's
                       util.store_infos through h5py): the HDF5 reader is tested against it on every machine.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import importlib.util

_spec = importlib.util.spec_from_file_location("_hdf5", os.path.join(ROOT, "vmc_pde_b200", "_hdf5.py"))
_hdf5 = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_hdf5)

REF = "/root/reference/vmc_fluids/paper_plot"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "ref_wiener_T10": "data_phaseSpace/Wiener/Nsamples10000_T10.0/infos.hdf5",
    "ref_wiener_Tdiff": "data_phaseSpace/Wiener/Nsamples10000_Tdifferent/infos.hdf5",
    "ref_inn_Tdiff": "data_phaseSpace/INN/NsamplesTDVP10000_NsamplesObs10000_Tdifferent/infos.hdf5",
    "ref_diff8_gauss": "data_diffusion/dim8_Gauss_NsamplesTDVP10000_NsamplesObs10000/infos.hdf5",
    "ref_diff8_student": "data_diffusion/dim8_StudentT_nu2_NsamplesTDVP10000_NsamplesObs10000/infos.hdf5",
}
SCALARS = ["times", "entropy", "integral_1sigma", "integral_0.5sigma", "integral_0.1sigma", "max_grad", "solver_res",
           "tdvp_error", "dist_params"]


def main():
    for name, rel in FILES.items():
        d = _hdf5.read(os.path.join(REF, rel))
        out = {"source": rel, "keys": np.array(sorted(d.keys())), "shapes": np.array([str(d[k].shape) for k in sorted(d.keys())])}
        n = len(d["times"])
        if name == "ref_wiener_T10":
            keep = np.arange(n)                                  # the whole deterministic trajectory
        elif name == "ref_wiener_Tdiff":
            keep = np.arange(1)
        else:
            keep = np.unique(np.concatenate([np.arange(0, 8), np.arange(8, n, 6), [n - 1]]))
        out["index"] = keep
        for k in SCALARS + ["x1", "covar", "x3", "x4", "x5", "x6"]:
            if k in d:
                out[k] = d[k][keep]
        if "ev" in d:                                            # spectra: first records, mid run, end
            sel = np.array([0, 1, n // 2, n - 1])
            out["ev_index"], out["ev"], out["snr"] = sel, d["ev"][sel], d["snr"][sel]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, n, "records ->", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")
    # one stored file as it is (the smallest, 184 KiB): a real h5py / libhdf5 product for the reader test that runs everywhere
    import shutil
    shutil.copyfile(os.path.join(REF, FILES["ref_wiener_Tdiff"]), os.path.join(HERE, "ref_wiener_Tdiff_infos.hdf5"))


if __name__ == "__main__":
    main()
