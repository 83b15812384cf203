"""Seeded synthetic workloads of the shapes BASELINE.json names.

Pure numpy, no compute path: the same arrays are handed to the CUDA path, to the oracle in
the tests, and to the CPU baseline in bench.py, so that all three see identical inputs.

Shapes (P test params, S GP samples, T time steps, D DOFs, decoder widths, activation):
  #1 examples/burgers1d.yml as shipped            441 x   20, T=1001, D= 1001, [5,100,1001] sigmoid
  #2 Burgers1D-shaped (the bench workload)         441 x 1000, T= 501, D= 1001, [5,100,1001] sigmoid
  #3 Burgers2D-shaped (legacy)                     441 x 1000, T= 201, D= 3600, [5,20,20,20,100,3600] softplus
  #4 Vlasov1D1V-shaped (legacy)                    441 x 2000, T= 251, D= 4096, [5,50,50,50,200,1000,4096] softplus
  #5 rising-bubble-shaped (8 GPU)                 4096 x 2000, T= 301, D=10000, [5,20,50,200,1000,10000] softplus
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


@dataclass
class Workload:
    name: str
    grid: tuple            # test grid (g0, g1); P = g0*g1, 'ij' order: first parameter slowest
    S: int
    T: int
    widths: list           # decoder widths [n_z, ..., D]
    act: str
    dt: float
    seed: int = 0
    n_z: int = 5

    @property
    def P(self) -> int:
        return int(self.grid[0] * self.grid[1])

    @property
    def D(self) -> int:
        return int(self.widths[-1])

    def decoder_flops_per_rollout(self) -> float:
        """Algorithmic decoder FLOP per rollout = T * 2 * sum_i L_i L_{i+1}."""
        L = self.widths
        return float(self.T) * 2.0 * sum(L[i] * L[i + 1] for i in range(len(L) - 1))


WORKLOADS = {
    "burgers1d_shipped": Workload("burgers1d_shipped", (21, 21), 20, 1001, [5, 100, 1001], "sigmoid", 1.0 / 1000),
    "burgers1d": Workload("burgers1d", (21, 21), 1000, 501, [5, 100, 1001], "sigmoid", 1.0 / 500),
    "burgers2d": Workload("burgers2d", (21, 21), 1000, 201, [5, 20, 20, 20, 100, 3600], "softplus", 1.0 / 200),
    "vlasov1d1v": Workload("vlasov1d1v", (21, 21), 2000, 251, [5, 50, 50, 50, 200, 1000, 4096], "softplus", 0.02),
    "rising_bubble": Workload("rising_bubble", (64, 64), 2000, 301, [5, 20, 50, 200, 1000, 10000], "softplus", 0.01),
    # small cases the oracle finishes in seconds
    "tiny": Workload("tiny", (3, 2), 8, 33, [5, 16, 64], "sigmoid", 1.0 / 32),
    "small": Workload("small", (4, 3), 20, 65, [5, 100, 301], "sigmoid", 1.0 / 64),
    "small_s200": Workload("small_s200", (3, 2), 200, 21, [5, 100, 1001], "sigmoid", 1.0 / 20),
    "small_softplus": Workload("small_softplus", (3, 2), 40, 17, [5, 20, 20, 100, 260], "softplus", 1.0 / 16),
    # ragged shapes: 2 DOF-tile groups (D > 1024), 2 sample chunks with a ragged tail, K not a multiple of 16
    "ragged": Workload("ragged", (3, 2), 130, 9, [5, 24, 1300], "tanh", 1.0 / 8),
    # decoder the fused tcgen05 kernel does not cover (K_last > 112) -> CUDA-core path
    "wide_k": Workload("wide_k", (3, 2), 24, 7, [5, 16, 160, 200], "softplus", 1.0 / 6),
    # same path through its 128 x 128 tiles (rows % 4 == 0, >= 64 samples): ragged sample / dof / k tails
    # the Vlasov-shaped decoder's tail at full width (K_last = 1000, D = 4096), few rows: 16 k-chunks x 3 split passes
    "wide_k1000": Workload("wide_k1000", (3, 2), 300, 4, [5, 50, 200, 1000, 4096], "softplus", 1.0 / 3),
    "wide_k128": Workload("wide_k128", (3, 2), 130, 5, [5, 40, 164, 300], "softplus", 1.0 / 4),
}


def make_decoder(widths, seed=0):
    """Xavier-uniform weights / default nn.Linear bias range (reference latent_space.py:179-189
    initialises with `xavier_uniform_`; biases keep nn.Linear's U(-1/sqrt(fan_in), +))."""
    rs = np.random.RandomState(1000 + seed)
    Ws, bs = [], []
    for i in range(len(widths) - 1):
        fi, fo = widths[i], widths[i + 1]
        a = np.sqrt(6.0 / (fi + fo))
        Ws.append(rs.uniform(-a, a, size=(fo, fi)).astype(np.float32))
        bb = 1.0 / np.sqrt(fi)
        bs.append(rs.uniform(-bb, bb, size=(fo,)).astype(np.float32))
    return Ws, bs


def make_inputs(wl: Workload, p_slice=None):
    """Returns dict(coefs fp64 [P,S,n(n+1)], z0 fp32 [P,n], t_grid fp64 [T], params fp64 [P,2]).

    Mean coefficients: A(p) = 3 (G - G^T) - 0.3 I + smooth variation over the grid, b ~ N(0,1);
    coefficient vector = row-major flatten of [b ; A^T] (reference sindy.py:72,80,112-113).
    Draw spread grows with the distance to the nearest of the 4 grid corners (the training
    points of the shipped example), mimicking the GP posterior; draws come from a seeded
    legacy RandomState in param-major (p, s, k) order."""
    n = wl.n_z
    g0, g1 = wl.grid
    a = np.linspace(0.0, 1.0, g0)
    w = np.linspace(0.0, 1.0, g1)
    A_, W_ = np.meshgrid(a, w, indexing="ij")          # first parameter slowest (param.py:212)
    params = np.stack([A_.ravel(), W_.ravel()], axis=1)
    P = params.shape[0]

    rs = np.random.RandomState(2000 + wl.seed)
    G = rs.normal(size=(n, n))
    A0 = 3.0 * (G - G.T) - 0.3 * np.eye(n)
    G1 = rs.normal(size=(n, n)); G2 = rs.normal(size=(n, n))
    A1 = 0.6 * (G1 - G1.T); A2 = 0.6 * (G2 - G2.T)
    b0 = rs.normal(size=n); b1 = 0.3 * rs.normal(size=n); b2 = 0.3 * rs.normal(size=n)
    z0 = rs.uniform(-1.0, 1.0, size=(P, n)).astype(np.float32)

    mean = np.zeros((P, n * (n + 1)))
    for p in range(P):
        x, y = params[p]
        A = A0 + x * A1 + y * A2
        b = b0 + x * b1 + y * b2
        R = np.concatenate([b[None, :], A.T], axis=0)   # (n+1, n): row 0 constant, rows 1.. = A^T
        mean[p] = R.ravel()

    corners = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float64)
    dist = np.min(np.linalg.norm(params[:, None, :] - corners[None, :, :], axis=2), axis=1)
    # a mild asymmetric bump makes the arg-max unique with a visible margin
    bump = 1.0 + 0.25 * np.exp(-((params[:, 0] - 0.55) ** 2 + (params[:, 1] - 0.40) ** 2) / 0.02)
    sigma = 0.04 * dist * bump                           # relative spread of each coefficient

    # p_slice: None (whole grid), a slice, or an explicit list of test-parameter indices
    sel = range(P) if p_slice is None else (range(*p_slice.indices(P)) if isinstance(p_slice, slice) else [int(p) for p in p_slice])
    coefs = np.zeros((len(sel), wl.S, n * (n + 1)))
    for o, p in enumerate(sel):
        rp = np.random.RandomState((3000 + wl.seed) * 100003 + p)
        std = sigma[p] * (np.abs(mean[p]) + 0.1)
        coefs[o] = rp.normal(mean[p][None, :], std[None, :], size=(wl.S, n * (n + 1)))
    idx = np.array(list(sel))
    t_grid = np.arange(wl.T, dtype=np.float64) * wl.dt
    return dict(coefs=coefs, z0=z0[idx], t_grid=t_grid, params=params[idx], mean=mean[idx], index=idx)


def make_noisy_series(B: int, T: int, n: int, dt: float, seed: int = 0, noise: float = 0.02) -> np.ndarray:
    """`make_latent_series` plus seeded white noise: the SINDy residual is then well above fp32 round-off, which
    the gradient tests of the calibration losses need (an exactly affine series has a pure-noise residual)."""
    Z = make_latent_series(B, T, n, dt, seed)
    rs = np.random.RandomState(9000 + seed)
    return (Z + noise * rs.normal(size=Z.shape)).astype(np.float32)


def make_latent_series(B: int, T: int, n: int, dt: float, seed: int = 0) -> np.ndarray:
    """Well-conditioned training series for calibration tests: trajectories of an oscillatory
    affine system (cond([1|Z]) ~ 10-100, SURVEY.md §7 'lstsq semantics').  fp32 [B, T, n]."""
    from scipy.linalg import expm
    rs = np.random.RandomState(4000 + seed)
    out = np.zeros((B, T, n), dtype=np.float32)
    for b in range(B):
        G = rs.normal(size=(n, n))
        A = 3.0 * (G - G.T) - 0.3 * np.eye(n)
        bb = rs.normal(size=n)
        M = np.zeros((n + 1, n + 1)); M[:n, :n] = A; M[:n, n] = bb
        E = expm(M * dt)
        z = np.concatenate([rs.uniform(-1, 1, size=n), [1.0]])
        for k in range(T):
            out[b, k] = z[:n]
            z = E @ z
    return out


def make_train_problem():
    """Tiny autoencoder-training problem shared by the fixture generator (tests/golden/gen_golden.py::gen_train, which
    runs the reference's own train()) and the GPU test: 4 training parameters, nt = 33, nx = 64, travelling smooth
    pulses as the FOM snapshots.  Returns (nt, nx, train_space, X fp32 [4, nt, nx], trainer config)."""
    nt, nx = 33, 64
    x = np.linspace(-3.0, 3.0, nx); t = np.linspace(0.0, 1.0, nt)
    train_space = np.array([[0.7, 0.9], [0.7, 1.1], [0.9, 0.9], [0.9, 1.1]])
    X = np.stack([a * np.exp(-(x[None, :] - 1.5 * t[:, None]) ** 2 / 2 / w / w) * np.cos(2 * t[:, None])
                  for a, w in train_space]).astype(np.float32)
    cfg = {"n_samples": 8, "lr": 1e-3, "n_iter": 6, "max_iter": 6, "max_greedy_iter": 1, "ld_weight": 0.1,
           "coef_weight": 1e-3}
    return nt, nx, train_space, X, cfg
