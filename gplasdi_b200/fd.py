"""Host-side mirror of the SBP finite-difference schemes (reference src/lasdi/fd.py:59-222).

The CUDA stencil kernel (csrc/calibrate.cu) applies these operators without ever forming a
matrix; this module only exposes the scheme registry (same keys as the reference's `FDdict`)
and a dense-matrix view for inspection / tests.
"""
from __future__ import annotations

import numpy as np

from .sbp_tables import SBP_TABLES, SBP_IDS


class SbpScheme:
    """One summation-by-parts first-derivative scheme (interior order 2w, boundary order w)."""

    def __init__(self, name: str):
        tab = SBP_TABLES[name]
        self.name = name
        self.id = SBP_IDS[name]
        self.halfwidth = tab["halfwidth"]
        self.leftBdrDepth = tab["depth"]
        self.leftBdrWidth = list(tab["widths"])
        self.leftBdrNorm = list(tab["norm"])
        self._tab = tab

    def min_length(self) -> int:
        return 2 * self.leftBdrDepth

    def dense(self, nt: int) -> np.ndarray:
        """(nt, nt) fp32 operator: antisymmetric interior band, dense left-boundary rows,
        right boundary = negated mirror (fd.py:32-39); values rounded to fp32 (fd.py:54)."""
        if nt < self.min_length():
            raise ValueError(f"{self.name} needs at least {self.min_length()} points")
        D = np.zeros((nt, nt))
        d, tab = self.leftBdrDepth, self._tab
        for r in range(d, nt - d):
            for k in range(1, self.halfwidth + 1):
                D[r, r + k] = tab["interior"][k - 1]
                D[r, r - k] = -tab["interior"][k - 1]
        for r in range(d):
            row = np.asarray(tab["rows"][r])
            D[r, : len(row)] = row
            D[nt - 1 - r, nt - len(row):] = -row[::-1]
        return D.astype(np.float32)

    def norm(self, nt: int) -> np.ndarray:
        """Diagonal SBP norm (quadrature weights), second return value of the reference's getOperators."""
        h = np.ones(nt)
        h[: self.leftBdrDepth] = self.leftBdrNorm
        h[nt - self.leftBdrDepth:] = self.leftBdrNorm[::-1]
        return h.astype(np.float32)


FDdict = {name: SbpScheme(name) for name in ("sbp12", "sbp24", "sbp36", "sbp48")}
