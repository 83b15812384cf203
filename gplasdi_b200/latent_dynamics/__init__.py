"""Latent-dynamics plugin interface; same contract as the reference's abstract base class
(reference src/lasdi/latent_dynamics/__init__.py:4-65): `dim`, `nt`, `ncoefs`, `calibrate`,
`simulate`, `sample`, `export`, `load`."""
from __future__ import annotations

import numpy as np


class LatentDynamics:
    dim = -1
    nt = -1
    ncoefs = -1

    def __init__(self, dim_, nt_):
        self.dim = int(dim_)
        self.nt = int(nt_)
        if self.dim <= 0 or self.nt <= 0:
            raise AssertionError("latent dimension and number of time steps must be positive")

    def calibrate(self, Z, dt, compute_loss=True, numpy=False):
        """Fit the dynamics coefficients to latent series Z ([nt, dim] or [n_train, nt, dim])."""
        raise RuntimeError("Abstract function LatentDynamics.calibrate!")

    def simulate(self, coefs, z0, t_grid):
        """Integrate one trajectory from z0 on t_grid; returns [nt, dim]."""
        raise RuntimeError("Abstract function LatentDynamics.simulate!")

    def sample(self, coefs_sample, z0_sample, t_grid):
        """Trajectories for N (coefficient, initial condition) pairs; returns [N, nt, dim].
        Subclasses with a batched integrator override this; the default loops like the reference."""
        if len(coefs_sample) != len(z0_sample):
            raise AssertionError("need one initial condition per coefficient sample")
        return np.stack([self.simulate(c, z, t_grid) for c, z in zip(coefs_sample, z0_sample)], axis=0)

    def export(self):
        return {"dim": self.dim, "ncoefs": self.ncoefs}

    def load(self, dict_):
        assert self.dim == dict_["dim"]
        assert self.ncoefs == dict_["ncoefs"]
