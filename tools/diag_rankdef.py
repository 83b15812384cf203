#!/usr/bin/env python
"""Rank-deficient calibration backward: GPU kernel vs the fp64 pseudo-inverse formula, term by term (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gplasdi_b200.engine import Engine
from oracle import reference_path as orc

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "calibrate_rankdef.npz"))
eng = Engine(0)
for kind in ["const_col", "collinear", "two_lost", "near"]:
    Zf = g[kind + "_Z"]; T = Zf.shape[1]; dt = 1 / (T - 1); n = 5
    coefs, ls, lc, info = eng.sindy_fit(Zf, dt, "sbp12", 1)
    Z = Zf[0].astype(np.float64)
    D = orc.fd_operator_dense(T, "sbp12").astype(np.float64)
    X = np.concatenate([np.ones((T, 1)), Z], 1); Y = D @ Z / dt
    Xp = np.linalg.pinv(X, rcond=float(np.finfo(np.float32).eps) * T); R = Xp @ Y; E = Y - X @ R
    Mp = Xp @ Xp.T; N = np.eye(n + 1) - Xp @ X
    Rg = coefs.cpu().numpy()[0].reshape(6, 5).astype(np.float64)
    print(kind, "info", int(info[0]), "coef err", np.abs(Rg - R).max(), "exact zeros in R (gpu/ref):", int((Rg == 0).sum()), int((np.abs(R) < 1e-12).sum()))
    for gs, gc in [(0.7, 0.0), (0.0, 1e-3), (0.7, 1e-3)]:
        c = 2 * gs / (T * n)
        Rbar = gc * np.sign(np.where(np.abs(R) < 1e-12, 0.0, R)) - c * X.T @ E
        Bm = Mp @ Rbar
        Ybar = X @ Bm + c * E
        Xbar = E @ Bm.T - Ybar @ R.T + X @ ((Mp @ R) @ (N @ Rbar).T)
        Zbar = Xbar[:, 1:] + D.T @ Ybar / dt
        got = eng.sindy_fit_backward(Zf, coefs, gs * torch.ones(1), gc * torch.ones(1), dt, "sbp12", 1).cpu().numpy()[0]
        T3 = (X @ ((Mp @ R) @ (N @ Rbar).T))[:, 1:]
        no3 = Zbar - T3
        T3t = (X @ ((N @ Rbar) @ (Mp @ R).T))[:, 1:]
        sc = np.abs(Zbar).max()
        d = got - no3
        print(f"      |got-no3| {np.abs(d).max()/sc:.2e} |T3| {np.abs(T3).max()/sc:.2e} |d-T3| {np.abs(d-T3).max()/sc:.2e} |d+T3| {np.abs(d+T3).max()/sc:.2e} "
              f"|d-T3t| {np.abs(d-T3t).max()/sc:.2e}  lstsq coef of d on T3: {float((d*T3).sum()/(T3*T3).sum()):.3f}")
        print(f"   gs {gs} gc {gc}: err {np.abs(got - Zbar).max() / np.abs(Zbar).max():.2e}   (vs formula without 3rd term {np.abs(got - no3).max() / np.abs(Zbar).max():.2e})")
