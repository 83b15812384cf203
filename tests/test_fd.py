"""Independent validation of the generated SBP tables and the host mirror gplasdi_b200/fd.py."""
import os

import numpy as np
import pytest

from gplasdi_b200.fd import FDdict

ORDERS = {"sbp12": (1, 2), "sbp24": (2, 4), "sbp36": (3, 6), "sbp48": (4, 8)}   # (boundary, interior)


@pytest.mark.parametrize("name", list(ORDERS))
def test_matches_reference_matrices(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "fd_operators.npz"))
    for nt in (17, 33):
        assert np.array_equal(FDdict[name].dense(nt), g[f"{name}_{nt}"])


@pytest.mark.parametrize("name", list(ORDERS))
def test_structure_and_accuracy(name):
    nt = 41
    D = FDdict[name].dense(nt).astype(np.float64)
    assert np.allclose(D.sum(axis=1), 0.0, atol=2e-6)              # derivative of a constant
    assert np.allclose(D[::-1, ::-1], -D)                           # centro-antisymmetric closure
    H = np.diag(FDdict[name].norm(nt).astype(np.float64))
    Q = H @ D
    B = np.zeros((nt, nt)); B[0, 0] = -1.0; B[-1, -1] = 1.0
    assert np.allclose(Q + Q.T, B, atol=5e-6)                       # summation-by-parts property
    x = np.linspace(0.0, 1.0, nt); h = x[1] - x[0]
    pb, pi = ORDERS[name]
    for deg in range(1, pb + 1):                                    # exact on the whole grid
        assert np.allclose(D @ x ** deg / h, deg * x ** (deg - 1), atol=2e-4)
    d = FDdict[name].leftBdrDepth
    for deg in range(1, pi + 1):                                    # exact in the interior
        err = (D @ x ** deg / h - deg * x ** (deg - 1))[d + 4: nt - d - 4]
        assert np.max(np.abs(err)) < 5e-4


def test_short_series_rejected():
    with pytest.raises(ValueError):
        FDdict["sbp48"].dense(15)
