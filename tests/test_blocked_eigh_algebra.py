"""CPU: the algebra the blocked eigensolver kernels implement (tools/proto/tridiag_blocked.py is their NumPy statement):
panel tridiagonalisation with a read-only trailing matrix, w at the next pivot row recomputed from the reduced dots,
w.v without a second reduction, masked 128-aligned trailing update, compact-WY back-transformation with dlarft's T."""
import importlib.util
import pathlib

import numpy as np
import pytest

_p = pathlib.Path(__file__).resolve().parents[1] / "tools" / "proto" / "tridiag_blocked.py"
_spec = importlib.util.spec_from_file_location("tridiag_blocked", _p)
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)


@pytest.mark.parametrize("n,nb,kb", [(1, 8, 8), (2, 8, 8), (3, 4, 4), (17, 8, 8), (64, 8, 16), (101, 16, 8), (130, 32, 32)])
def test_blocked_tridiagonalisation_and_backtransform(n, nb, kb):
    rng = np.random.default_rng(n)
    B = rng.normal(size=(n, n)); A = (B + B.T) / 2
    if n >= 17:
        A[3, 5:] = 0; A[5:, 3] = 0            # a reflector-free column (sigma == 0)
    d, e, tau, Vst, _ = proto.tridiag_blocked(A, nb=nb)
    T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
    lam, Z = np.linalg.eigh(T)
    ref = np.linalg.eigvalsh(A)
    V = proto.backtransform_blocked(Z, Vst, tau, kb=kb)
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(lam - ref).max() < 1e-12 * scale
    assert np.abs(A @ V - V * lam).max() < 1e-12 * scale
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12


def test_panel_gram_byproduct_is_the_reflector_gram():
    """The dots V^T v the panel kernel reduces anyway are the strictly upper part of Y^T Y inside a panel."""
    rng = np.random.default_rng(0)
    n, nb = 40, 8
    B = rng.normal(size=(n, n)); A = (B + B.T) / 2
    _, _, _, Vst, Gst = proto.tridiag_blocked(A, nb=nb)
    G = Vst @ Vst.T
    for j in range(n - 2):
        j0 = j // nb * nb
        assert np.allclose(Gst[j, j0:j], G[j, j0:j], atol=1e-13)
