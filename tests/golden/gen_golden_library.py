#!/usr/bin/env python
"""Golden trajectories of the polynomial / sine SINDy libraries, produced by RUNNING THE REFERENCE'S OWN
`simulate_sindy` (legacy/Vlasov1D1V/utils/utils_sindy.py:91-170; build container only):

    python tests/golden/gen_golden_library.py      ->  tests/golden/simulate_library.npz

For each of the four libraries (poly_order 1/2 x sine_term False/True) and latent dimensions 3 and 5 it stores the
seeded coefficient matrices `coef_<tag>` [N, |Theta|, n], initial conditions `z0_<tag>` [N, n] (fp32 values, as the
encoder produces them), the time grid and `Z_<tag>` [N, nt, n] = what `simulate_sindy` returned.
Coefficients: a damped rotation (the affine part of gplasdi_b200.synth) plus small quadratic and sine couplings, so
the quadratic systems stay bounded over the grid.
"""
import contextlib, io, os, sys
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/legacy/Vlasov1D1V")
sys.path.insert(0, "/root/reference/legacy/Vlasov1D1V/utils")

import numpy as np
np.Inf = np.inf
import utils_sindy as ref   # the reference


def make_case(n, poly_order, sine_term, N, seed):
    rs = np.random.RandomState(7000 + seed)
    n_q = ref.lib_size(n, poly_order) if poly_order == 2 else 0
    terms = 1 + n + n_q + (n if sine_term else 0)
    coefs = np.zeros((N, terms, n))
    for k in range(N):
        G = rs.normal(size=(n, n))
        A = 2.0 * (G - G.T) - 0.4 * np.eye(n)
        coefs[k, 0] = 0.5 * rs.normal(size=n)
        coefs[k, 1:1 + n] = A.T
        if n_q:
            coefs[k, 1 + n:1 + n + n_q] = 0.08 * rs.normal(size=(n_q, n))
        if sine_term:
            coefs[k, 1 + n + n_q:] = 0.7 * rs.normal(size=(n, n))
    z0 = rs.uniform(-1.0, 1.0, size=(N, n)).astype(np.float32)
    return coefs, z0, n_q


def main():
    out = {}
    t_grid = np.linspace(0.0, 2.0, 101)
    out["t_grid"] = t_grid
    tags = []
    for n in (3, 5):
        for poly_order in (1, 2):
            for sine_term in (False, True):
                tag = f"n{n}_p{poly_order}_s{int(sine_term)}"
                coefs, z0, n_q = make_case(n, poly_order, sine_term, 6, seed=len(tags))
                with contextlib.redirect_stdout(io.StringIO()):      # the poly-1 + sine branch prints inside its RHS
                    Z = ref.simulate_sindy([c for c in coefs], z0.astype(np.float64), t_grid, n, poly_order=poly_order,
                                           library_size=n_q if poly_order == 2 else None, sine_term=sine_term)
                assert np.all(np.isfinite(Z)) and np.max(np.abs(Z)) < 50.0, (tag, np.max(np.abs(Z)))
                out[f"coef_{tag}"] = coefs; out[f"z0_{tag}"] = z0; out[f"Z_{tag}"] = Z
                tags.append(tag)
                print(tag, Z.shape, "max|Z|", float(np.max(np.abs(Z))))
    out["tags"] = np.array(tags)
    np.savez_compressed(os.path.join(HERE, "simulate_library.npz"), **out)


if __name__ == "__main__":
    main()
