"""Gaussian-process interpolation of the SINDy coefficients over parameter space; same functions
as the reference module (reference src/lasdi/gp.py:5-80).

The GP FIT (hyper-parameter optimisation, `fit_gps`) stays on scikit-learn on the host.  The posterior and
the draws have two implementations with the same results:
  * host: `eval_gp` / `sample_coefs` / `sample_coefs_all` -- scikit-learn `predict` and ONE vectorised
    `np.random.normal` call per parameter, consuming numpy's global legacy MT19937 stream exactly like the
    reference's scalar double loop (gp.py:76-78);
  * device: `eval_gp_device` / `sample_coefs_all_device` -- `glasdi_gp_predict` and
    `glasdi_normal_mt19937`, which reproduces the same MT19937 / polar Box-Muller stream on the GPU and hands
    numpy's global generator back in the state the reference would leave it in.  The coefficient draws
    ([P, S, n_coef] fp64, 106 MB at the bench shape) are then born in HBM and never cross PCIe.
Seeded runs select the same parameter either way.
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


def gp_arrays(gp_dictionnary):
    """The fitted quantities `predict` uses, stacked over the GPs (scikit-learn attributes X_train_, alpha_, L_,
    kernel_ = ConstantKernel(c) * RBF(l), _y_train_mean/_std)."""
    g0 = gp_dictionnary[0]
    X = np.asarray(g0.X_train_, dtype=np.float64)
    alpha, L, amp, ls, ym, ys = [], [], [], [], [], []
    for gp in gp_dictionnary:
        k = gp.kernel_
        if not (isinstance(k.k1, ConstantKernel) and isinstance(k.k2, RBF) and np.ndim(k.k2.length_scale) == 0):
            raise NotImplementedError("device GP posterior supports ConstantKernel * isotropic RBF (what fit_gps builds)")
        if gp.X_train_.shape != X.shape or not np.array_equal(gp.X_train_, X):
            raise ValueError("all coefficient GPs must share their training inputs")
        alpha.append(np.asarray(gp.alpha_, dtype=np.float64).reshape(-1))
        L.append(np.asarray(gp.L_, dtype=np.float64))
        amp.append(float(k.k1.constant_value)); ls.append(float(k.k2.length_scale))
        ym.append(float(np.ravel(gp._y_train_mean)[0])); ys.append(float(np.ravel(gp._y_train_std)[0]))
    return X, np.stack(alpha), np.stack(L), np.array(amp), np.array(ls), np.array(ym), np.array(ys)


def eval_gp_device(gp_dictionnary, param_grid, engine=None):
    """`eval_gp` on the GPU: (pred_mean, pred_std) as CUDA fp64 tensors [n_points, n_coef]."""
    from .engine import default_engine
    eng = engine or default_engine()
    param_grid = np.asarray(param_grid, dtype=np.float64)
    if param_grid.ndim == 1:
        param_grid = param_grid.reshape(1, -1)
    X, alpha, L, amp, ls, ym, ys = gp_arrays(gp_dictionnary)
    return eng.gp_predict(X, alpha, L, amp, ls, param_grid, ym, ys)


def sample_coefs_all_device(gp_dictionnary, param_grid, n_samples, engine=None):
    """`sample_coefs_all` on the GPU: CUDA fp64 [n_points, n_samples, n_coef], drawn from numpy's global legacy
    stream (advanced exactly as the reference's loops would advance it)."""
    from .engine import default_engine
    eng = engine or default_engine()
    mean, std = eval_gp_device(gp_dictionnary, param_grid, eng)
    out, _ = eng.normal_mt19937(mean, std, n_samples)
    return out
