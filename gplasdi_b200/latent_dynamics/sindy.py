"""SINDy latent dynamics on the B200 path; drop-in for the reference class of the same name
(reference src/lasdi/latent_dynamics/sindy.py:8-123).

  calibrate -> glasdi_sindy_fit   (SBP stencil dZ/dt + fp64 normal-equation solve per series)
  simulate  -> glasdi_rollout     (fp64 exact affine propagator; the reference calls scipy odeint)
  sample    -> glasdi_rollout over the whole batch in one launch

Signatures, return types and the coefficient layout (row-major flatten of the (dim+1, dim)
least-squares solution, row 0 = constant term) are the reference's.  When `Z` carries an autograd
graph (the reference's training loop backpropagates both losses into the encoder, gplasdi.py:186-197)
the two losses returned by `calibrate` are differentiable: `_SindyFit` pairs the forward kernel with
the analytic backward kernel `glasdi_sindy_fit_backward`; the coefficients are returned detached,
as in the reference (sindy.py:80).
"""
from __future__ import annotations

import numpy as np
import torch

from . import LatentDynamics
from ..fd import FDdict
from ..engine import default_engine
from .. import _lib


def _is_uniform(t_grid) -> bool:
    t = np.asarray(t_grid, dtype=np.float64)
    if t.ndim != 1 or t.size < 1:
        raise ValueError("t_grid must be a 1-d array")
    if t.size <= 2:
        return True
    d = np.diff(t)
    return bool(np.allclose(d, d[0], rtol=1e-9, atol=1e-14))


def _uniform_dt(t_grid) -> float:
    """Step of a uniform grid (the fused greedy step and the exact affine propagator need one; `simulate` / `sample` take
    any grid through glasdi_rollout_grid)."""
    t = np.asarray(t_grid, dtype=np.float64)
    if not _is_uniform(t):
        raise ValueError("this call needs a uniform time grid (the reference's physics grids are uniform); "
                         "SINDy.simulate / sample accept arbitrary grids")
    return 0.0 if t.size == 1 else float(t[1] - t[0])


def lib_size(n_z: int, poly_order: int) -> int:
    """Number of polynomial terms of order 2..poly_order (legacy/Vlasov1D1V/utils/utils_sindy.py:12-27)."""
    from math import comb
    return sum(comb(n_z + k - 1, k) for k in range(poly_order + 1)) - n_z - 1


def simulate_sindy(sindy_coef, Z0, t_grid, n_z, poly_order=1, library_size=None, sine_term=False, engine=None):
    """Drop-in for the legacy drivers' `simulate_sindy` (legacy/Vlasov1D1V/utils/utils_sindy.py:91-170): integrates
    dz/dt = Theta(z) sindy_coef[i] from Z0[i] for every i, Theta = [1, z, z_i z_j (i <= j) if poly_order == 2, sin z if
    sine_term].  sindy_coef: sequence of [|Theta|, n_z] matrices; returns np.ndarray fp64 [len(sindy_coef), nt, n_z].
    All systems run in ONE launch of the CUDA rollout (`rollout_lib_kernel`; the affine case on the exact propagator)."""
    eng = engine if engine is not None else default_engine()
    c = np.stack([np.asarray(ci, dtype=np.float64) for ci in sindy_coef])
    N = c.shape[0]
    n_q = lib_size(n_z, 2) if poly_order == 2 else 0
    if library_size is not None and poly_order == 2 and int(library_size) != n_q:
        raise ValueError(f"library_size {library_size} does not match lib_size({n_z}, 2) = {n_q}")
    terms = 1 + n_z + n_q + (n_z if sine_term else 0)
    if c.shape[1:] != (terms, n_z):
        raise ValueError(f"sindy_coef matrices must be [{terms}, {n_z}] for this library, got {c.shape[1:]}")
    z = np.asarray(Z0, dtype=np.float32).reshape(N, n_z)
    prev = eng.get_option(_lib.OPT_SINDY_LIBRARY)
    eng.set_library(poly_order, sine_term)
    try:
        Z = eng.rollout(c.reshape(N, 1, terms * n_z), z, len(t_grid), _uniform_dt(t_grid))
        eng.check_numeric()          # a blown-up trajectory (RK4 step bound exceeded / non-finite) raises
    finally:
        eng.set_option(_lib.OPT_SINDY_LIBRARY, prev)
    return Z[:, 0].cpu().numpy()


class _SindyFit(torch.autograd.Function):
    """(loss_sindy[B], loss_coef[B], coefs[B, n(n+1)], info[B]) of a batch of series on the CUDA kernels; the
    backward is one launch of the analytic gradient kernel (no autograd graph through the solve)."""

    @staticmethod
    def forward(ctx, Z, engine, dt, fd_type, norm_order):
        Zc = Z.detach().to(device=engine.tdev, dtype=torch.float32).contiguous()
        coefs, ls, lc, info = engine.sindy_fit(Zc, dt, fd_type, norm_order)
        ctx.save_for_backward(Zc, coefs)
        ctx.engine, ctx.dt, ctx.fd_type, ctx.norm_order = engine, dt, fd_type, norm_order
        ctx.in_device, ctx.in_dtype = Z.device, Z.dtype
        ctx.mark_non_differentiable(coefs, info)
        return ls, lc, coefs, info

    @staticmethod
    def backward(ctx, g_ls, g_lc, _g_coefs, _g_info):
        Zc, coefs = ctx.saved_tensors
        gZ = ctx.engine.sindy_fit_backward(Zc, coefs, g_ls, g_lc, ctx.dt, ctx.fd_type, ctx.norm_order)
        return gZ.to(device=ctx.in_device, dtype=ctx.in_dtype), None, None, None, None


class SINDy(LatentDynamics):
    fd_type = ""

    def __init__(self, dim, nt, config, engine=None):
        super().__init__(dim, nt)
        self.ncoefs = self.dim * (self.dim + 1)
        assert "sindy" in config
        opts = config["sindy"] or {}
        self.fd_type = opts.get("fd_type", "sbp12")
        if self.fd_type not in FDdict:
            raise KeyError(f"unknown fd_type {self.fd_type!r}; choose from {sorted(FDdict)}")
        self.fd = FDdict[self.fd_type]
        if self.nt < self.fd.min_length():
            raise AssertionError(f"{self.fd_type} needs nt >= {self.fd.min_length()}")
        self.coef_norm_order = opts.get("coef_norm_order", 1)
        if self.coef_norm_order not in (1, 2, "fro"):
            raise ValueError("coef_norm_order must be 1, 2 or 'fro'")
        self._engine = engine

    @property
    def engine(self):
        if self._engine is None:
            self._engine = default_engine()
        return self._engine

    # -------------------------------------------------------------------------------------
    def compute_time_derivative(self, Z, Dt):
        """(1/Dt) * D_fd Z for one series [nt, dim] (or a batch [B, nt, dim]); returns a tensor on
        the input's device.  Reference sindy.py:89-102."""
        Zt = torch.as_tensor(Z)
        out = self.engine.fd_dzdt(Zt.detach(), Dt, self.fd_type)
        return out.to(Zt.device)

    def calibrate(self, Z, dt, compute_loss=True, numpy=False):
        """Reference sindy.py:41-87.  Z: torch tensor [nt, dim] or [n_train, nt, dim].
        Returns coefs (and the two losses, summed over the training series)."""
        Zt = torch.as_tensor(Z)
        batched = Zt.dim() == 3
        if not batched:
            assert Zt.dim() == 2
        Zb = Zt if batched else Zt.unsqueeze(0)
        assert Zb.shape[2] == self.dim          # (the reference does not check nt either)
        if compute_loss and torch.is_grad_enabled() and Zb.requires_grad:
            ls, lc, coefs, info = _SindyFit.apply(Zb, self.engine, float(dt), self.fd_type, self.coef_norm_order)
        else:
            coefs, ls, lc, info = self.engine.sindy_fit(Zb.detach(), dt, self.fd_type, self.coef_norm_order)
        # `info` = number of directions of the library matrix [1 | Z] the fit truncated, by LAPACK gelsy's own rank rule
        # (rcond = eps_fp32 * max(nt, dim + 1), what torch.linalg.lstsq applies at sindy.py:72): 0 = full rank; 1..dim = the
        # MINIMUM-NORM solution of the rank-truncated problem was returned, finite, as the reference does for the constant /
        # collinear latent columns of an untrained encoder; dim + 1 = the series is not finite.  The host copy of the
        # coefficients below synchronises anyway: read the status with it.
        info_h = info.cpu()
        self.last_truncated_rank = info_h.numpy().copy()
        bad = torch.nonzero(info_h > self.dim).flatten().tolist()
        if bad:
            raise RuntimeError(f"SINDy.calibrate: training series {bad} contain non-finite latent values")
        if batched:
            out = coefs.double().cpu().numpy() if numpy else coefs.cpu()
        else:
            out = coefs[0].cpu().numpy() if numpy else coefs[0].cpu()
        if not compute_loss:
            return out
        # losses live where Z lives (the reference computes them on the CPU copy of Z, gplasdi.py:183)
        return out, ls.sum().to(Zt.device), lc.sum().to(Zt.device)

    def simulate(self, coefs, z0, t_grid):
        """One trajectory; returns np.ndarray fp64 [nt, dim] like scipy odeint does in the reference."""
        c = np.asarray(coefs, dtype=np.float64).reshape(1, 1, self.ncoefs)
        z = np.asarray(z0, dtype=np.float32).reshape(1, self.dim)
        if _is_uniform(t_grid):
            Z = self.engine.rollout(c, z, len(t_grid), _uniform_dt(t_grid))
        else:                                # odeint takes any grid (sindy.py:115): step-adaptive RK4 kernel
            Z = self.engine.rollout_grid(c, z, t_grid)
        return Z[0, 0].cpu().numpy()

    def sample(self, coefs_sample, z0_sample, t_grid):
        """N trajectories in one launch; returns np.ndarray fp64 [N, nt, dim]."""
        c = np.asarray(coefs_sample, dtype=np.float64)
        z = np.asarray(z0_sample, dtype=np.float32)
        assert len(c) == len(z)
        N = c.shape[0]
        if _is_uniform(t_grid):
            Z = self.engine.rollout(c.reshape(N, 1, self.ncoefs), z.reshape(N, self.dim), len(t_grid), _uniform_dt(t_grid))
        else:
            Z = self.engine.rollout_grid(c.reshape(N, 1, self.ncoefs), z.reshape(N, self.dim), t_grid)
        return Z[:, 0].cpu().numpy()

    def rollout_ensemble(self, coef_samples, Z0, t_grid):
        """[P, S, ncoefs] draws and [P, dim] initial conditions -> CUDA tensor fp64 [P, S, nt, dim]
        (the array the reference fills in gplasdi.py:262-266)."""
        dt = _uniform_dt(t_grid)
        return self.engine.rollout(coef_samples, Z0, len(t_grid), dt)

    def export(self):
        d = super().export()
        d["fd_type"] = self.fd_type
        d["coef_norm_order"] = self.coef_norm_order
        return d
