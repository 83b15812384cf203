"""Gaussian-process interpolation of the SINDy coefficients over parameter space; same functions
as the reference module (reference src/lasdi/gp.py:5-80).

The GP fit / posterior stay on scikit-learn on the host in this round (the
device port is the next tier).  What changes is `sample_coefs_all`, which draws the coefficients
for every test parameter from the posterior moments with ONE vectorised call per parameter that
consumes numpy's global legacy MT19937 stream exactly like the reference's scalar double loop
(gp.py:76-78), so seeded runs select the same parameter.
"""
from __future__ import annotations

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import ConstantKernel, RBF


def fit_gps(X, Y):
    """One independent GP per coefficient: ConstantKernel * RBF(length_scale_bounds=(0.1, 1e5)),
    10 optimiser restarts, random_state=1 (reference gp.py:5-34).
    X [n_train, n_param], Y [n_train, n_coef] -> list of fitted regressors."""
    X = np.asarray(X)
    Y = np.asarray(Y)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    targets = [Y] if Y.ndim == 1 else list(Y.T)
    gps = []
    for yk in targets:
        kernel = ConstantKernel() * RBF(length_scale_bounds=(0.1, 1e5))
        gp = GaussianProcessRegressor(kernel=kernel, n_restarts_optimizer=10, random_state=1)
        gp.fit(X, yk)
        gps.append(gp)
    return gps


def eval_gp(gp_dictionnary, param_grid):
    """Posterior mean / std of every coefficient at every grid point (reference gp.py:36-53).
    Returns (pred_mean, pred_std), both [n_points, n_coef]."""
    param_grid = np.asarray(param_grid)
    if param_grid.ndim == 1:
        param_grid = param_grid.reshape(1, -1)
    n_points, n_coef = param_grid.shape[0], len(gp_dictionnary)
    pred_mean = np.zeros([n_points, n_coef])
    pred_std = np.zeros([n_points, n_coef])
    for k, gp in enumerate(gp_dictionnary):
        pred_mean[:, k], pred_std[:, k] = gp.predict(param_grid, return_std=True)
    return pred_mean, pred_std


def sample_coefs(gp_dictionnary, param, n_samples):
    """[n_samples, n_coef] draws for ONE parameter point (reference gp.py:55-80)."""
    param = np.asarray(param)
    if param.ndim == 1:
        param = param.reshape(1, -1)
    assert param.shape[0] == 1
    pred_mean, pred_std = eval_gp(gp_dictionnary, param)
    return np.random.normal(pred_mean[0][None, :], pred_std[0][None, :], size=(n_samples, len(gp_dictionnary)))


def sample_coefs_all(gp_dictionnary, param_grid, n_samples):
    """Draws for every grid point, [n_points, n_samples, n_coef], in the order the reference's
    per-parameter loop consumes the RNG (gplasdi.py:260: parameter-major, then sample, then coef).

    Note on the posterior: the reference evaluates `gp.predict` one point at a time; evaluated in
    one batch the moments agree to fp64 round-off (~1e-13 relative), which is what is used here."""
    param_grid = np.asarray(param_grid)
    if param_grid.ndim == 1:
        param_grid = param_grid.reshape(1, -1)
    pred_mean, pred_std = eval_gp(gp_dictionnary, param_grid)
    n_points, n_coef = pred_mean.shape
    out = np.empty([n_points, n_samples, n_coef])
    for i in range(n_points):
        out[i] = np.random.normal(pred_mean[i][None, :], pred_std[i][None, :], size=(n_samples, n_coef))
    return out
