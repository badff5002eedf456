"""In-tree build of libvmcpde.so (sm_100a) with plain nvcc; objects are compiled in parallel.

    python -m vmc_pde_b200.build [--force]

The library is written next to this file so that it travels with the repository snapshot to the GPU
box (it is git-ignored).  No CPU fallback is ever built into it.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvmcpde.so")
DIMS = (2, 3, 4, 5, 6, 8, 10, 12)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("VMCPDE_NVCC_EXTRA", "").split()
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"] + EXTRA
PLAIN = ["capi.cu", "gram.cu", "gram_split.cu", "reduce_kernels.cu", "solve.cu", "eigh.cu", "eigh_blocked.cu", "observables.cu", "particles.cu", "collectives.cu"]


def _units():
    units = []
    for f in PLAIN:
        if os.path.exists(os.path.join(CSRC, f)):
            units.append((f, [], f.replace(".cu", ".o")))
    for d in DIMS:        # largest first: the d = 12 multi-layer unit is the longest compile
        for ml in (1, 0):  # without / with the generic multi-layer SingleTrafo path (flow_core.cuh: VMC_ML)
            units.append(("flow_kernels_dim.cu", [f"-DVMC_DIM={d}", f"-DVMC_ML={ml}"], f"flow_kernels_dim{d}_ml{ml}.o"))
    return units


def _stamp():
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cc")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(unit):
    src, defs, obj = unit
    cmd = [NVCC, *FLAGS, *defs, "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    units = _units()
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(_compile, units))
    for obj, rc, out in results:
        if rc != 0:
            raise RuntimeError(f"nvcc failed for {obj}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
    objs = [os.path.join(OBJ, u[2]) for u in units]
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    open(stamp_file, "w").write(stamp)
    build_xla_shim(verbose)
    return LIB


def xla_ffi_include_dir():
    """Directory of xla/ffi/api/ffi.h when JAX is importable (probed on every build), else None."""
    try:
        import jax
        d = jax.ffi.include_dir()
        return d if os.path.exists(os.path.join(d, "xla", "ffi", "api", "ffi.h")) else None
    except Exception:
        return None


def build_xla_shim(verbose=True):
    """libvmcpde_xla.so: the XLA-FFI handlers of csrc/xla_ffi_shim.cc (north star: thin jax.ffi custom calls).  Built only
    when the XLA FFI headers exist; this image has no JAX, so here the function reports and returns None."""
    inc = xla_ffi_include_dir()
    if inc is None:
        if verbose:
            print("xla_ffi_shim.cc not compiled: `import jax` / jax.ffi.include_dir() unavailable in this environment", file=sys.stderr)
        return None
    out = os.path.join(HERE, "libvmcpde_xla.so")
    cmd = [NVCC, "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-I" + inc, "-I" + os.path.join(HERE, "..", "include"),
           os.path.join(CSRC, "xla_ffi_shim.cc"), "-L" + HERE, "-lvmcpde", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("XLA-FFI shim failed to compile:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
