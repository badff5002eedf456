"""CPU: the C-ABI library loads and exports every symbol include/vmcpde.h declares; product hygiene checks."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vmc_pde_b200 import build, _lib
    build.build(verbose=False)
    return _lib.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vmcpde.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vmcpde_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vmcpde.h but not exported by libvmcpde.so"


def test_loader_signatures_cover_header(lib):
    from vmc_pde_b200 import _lib
    assert set(declared_symbols()) == set(_lib.SIGNATURES), set(declared_symbols()) ^ set(_lib.SIGNATURES)


def test_host_only_entry_points(lib):
    """Calls that need no GPU: version, padding rule, flow handle + flat layout offsets (var_state.py:106-108)."""
    from vmc_pde_b200 import _capi, _lib
    assert lib.vmcpde_version() == 100
    assert [lib.vmcpde_padded_params(p) for p in (1, 37, 128, 129, 8187, 16385)] == [128, 128, 128, 256, 8192, 16512]
    # packed upper 128 x 128 tiles of a Pp x Pp Gram matrix (what crosses NVLink): 2080 tiles = 272 MB at Pp = 8192
    assert [lib.vmcpde_packed_tiles_len(p) for p in (128, 256, 8192)] == [16384, 3 * 16384, 2080 * 16384]
    ups = [[1, 0, 3][:2], [2, 3]]
    cfg, keep = _capi.make_flow_config(4, 2, (5,), "different_add", "Student_t", [[0, 1], [2, 3]], [[2, 3], [0, 1]], np.zeros(4))
    h = ctypes.c_void_p()
    _lib.check(lib.vmcpde_flow_create(ctypes.byref(cfg), ctypes.byref(h)))
    T = 5 + 2 * 5 + 2 + 5 * 2
    assert lib.vmcpde_flow_num_params(h) == 6 + 4 + 1 + 4 + 2 * 4 * T
    off = (ctypes.c_int32 * 6)()
    _lib.check(lib.vmcpde_flow_param_offsets(h, off))
    assert list(off) == [0, 6, 10, 11, 15, 15 + 4 * T]
    lib.vmcpde_flow_destroy(h)
    # error reporting
    bad, keep2 = _capi.make_flow_config(4, 1, (3,), "no_add", "Gauss", [[0, 0]], [[1, 2]], np.zeros(4))
    assert lib.vmcpde_flow_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"partition" in lib.vmcpde_last_error()
    two, keep3 = _capi.make_flow_config(4, 1, (3, 5), "no_add", "Gauss", [[0, 1]], [[2, 3]], np.zeros(4))
    _lib.check(lib.vmcpde_flow_create(ctypes.byref(two), ctypes.byref(h)))       # several hidden layers: net.py:53-58 loops over `intmediate`
    assert lib.vmcpde_flow_num_params(h) == 6 + 4 + 4 + 2 * ((3 + 2 * 3) + (5 + 3 * 5) + (2 + 5 * 2))
    lib.vmcpde_flow_destroy(h)
    gcf, keep6 = _capi.make_flow_config(4, 2, (5,), "different_add", "Student_t", [[0, 1], [2, 3]], [[2, 3], [0, 1]], np.zeros(4), global_change=True)
    _lib.check(lib.vmcpde_flow_create(ctypes.byref(gcf), ctypes.byref(h)))       # SingleBlock.global_change: + dim + 1 parameters per block
    assert lib.vmcpde_flow_num_params(h) == 6 + 4 + 1 + 4 + 2 * (4 * T + 5)
    _lib.check(lib.vmcpde_flow_param_offsets(h, off))
    assert list(off) == [0, 6, 10, 11, 15, 15 + 4 * T + 5]
    lib.vmcpde_flow_destroy(h)
    badv, keep7 = _capi.make_flow_config(4, 1, (3,), 0x204, "Gauss", [[0, 1]], [[2, 3]], np.zeros(4))
    assert lib.vmcpde_flow_create(ctypes.byref(badv), ctypes.byref(h)) != 0
    four, keep4 = _capi.make_flow_config(4, 1, (3, 3, 3, 3), "no_add", "Gauss", [[0, 1]], [[2, 3]], np.zeros(4))
    assert lib.vmcpde_flow_create(ctypes.byref(four), ctypes.byref(h)) == 2  # VMCPDE_EUNSUPPORTED: at most three hidden layers
    wide, keep5 = _capi.make_flow_config(4, 1, (3, 40), "no_add", "Gauss", [[0, 1]], [[2, 3]], np.zeros(4))
    assert lib.vmcpde_flow_create(ctypes.byref(wide), ctypes.byref(h)) == 2  # ... of at most 32 units each on the generic path


def test_string_sorted_block_order(lib):
    """depth > 10: flax sorts 'blocks_10' before 'blocks_2' (SURVEY appendix B)."""
    from vmc_pde_b200 import _capi, _lib
    depth = 12
    cfg, keep = _capi.make_flow_config(2, depth, (1,), "no_add", "Gauss", [[0]] * depth, [[1]] * depth, np.zeros(2))
    h = ctypes.c_void_p()
    _lib.check(lib.vmcpde_flow_create(ctypes.byref(cfg), ctypes.byref(h)))
    off = (ctypes.c_int32 * (4 + depth))()
    _lib.check(lib.vmcpde_flow_param_offsets(h, off))
    blocks = list(off)[4:]
    order = sorted(range(depth), key=lambda b: blocks[b])
    assert order == sorted(range(depth), key=lambda b: f"blocks_{b}")
    lib.vmcpde_flow_destroy(h)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vmc_pde_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "hostsim" not in txt or f == "flow_core.cuh" or f == "dc_core.cuh", f


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vmc_pde_b200 import sampler, var_state, _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_cuda()
    s = None
    with pytest.raises(RuntimeError, match="CUDA"):
        s = sampler.Sampler(dim=2, name="Gauss", mcmc_info={"offset": np.zeros(2)})
        var_state.VarState(s, 2, 1, 4, network_args={"intmediate": (1,), "offset": np.zeros(2), "latentSpaceName": "Gauss", "dim": 2})


def test_loader_fails_loudly_when_library_missing(tmp_path, monkeypatch):
    from vmc_pde_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libvmcpde.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


@pytest.mark.gpu
def test_torch_free_c_caller():
    """tests/cabi_smoke.cu: a plain C++/CUDA program (cudaMalloc, default stream, no torch, no Python) drives one right-hand side
    through the C-ABI -- sample -> local terms -> moments -> centring -> Gram (DMMA and tcgen05 split) -> eigh -> solve tail --
    and the two-call eigensolver; "raw pointers + stream, caller-owned workspaces" demonstrated, not asserted."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cabi_smoke")
    if not os.path.exists(exe):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe, exe + ".cu",
                               "-L" + os.path.join(root, "vmc_pde_b200"), "-lvmcpde", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../vmc_pde_b200"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CABI_SMOKE OK" in r.stdout, r.stdout + r.stderr
