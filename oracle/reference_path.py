"""CPU ORACLE for the GPLaSDI greedy-sampling / SINDy-calibration hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import it; nothing under
gplasdi_b200/ does (tests/test_boundary.py enforces that).

It restates the reference's algorithm for the path with the same third-party numerics the
reference itself calls (the reference has no native code and no own arithmetic: torch CPU
`Linear`/`linalg.lstsq`, scipy `odeint`, numpy `std`, numpy legacy RNG), so that it can run
on the GPU box where /root/reference does not exist.

Parity pinning: the reference ships NO tests / golden vectors for this path (SURVEY.md §4,
§8c), so the oracle is pinned against outputs of the reference itself, generated in the
build container by tests/golden/gen_golden.py (which imports /root/reference/src/lasdi) and
committed as tests/golden/*.npz; tests/test_oracle_golden.py checks every function below
against those fixtures (bit-exact where the library call sequence is identical).

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.integrate import odeint

from gplasdi_b200.sbp_tables import SBP_TABLES  # numeric constants only (no compute)

# ----------------------------------------------------------------------------------------
# fd.py: SBP first-derivative operator
# ----------------------------------------------------------------------------------------

def fd_operator_dense(nt: int, fd_type: str = "sbp12") -> np.ndarray:
    """Dense (nt x nt) fp32 SBP operator.

    Follows reference src/lasdi/fd.py:14-44 (`Stencil.getOperators`, non-periodic branch
    :32-39) and the fp32 cast of the values in `convert` (:46-57): interior rows carry the
    antisymmetric interior stencil, the first `depth` rows the dense left-boundary rows,
    the last `depth` rows the negated mirror image.
    """
    tab = SBP_TABLES[fd_type]
    hw, depth = tab["halfwidth"], tab["depth"]
    assert nt >= 2 * depth, "series too short for the boundary closure"
    D = np.zeros((nt, nt), dtype=np.float64)
    for k in range(1, hw + 1):
        c = tab["interior"][k - 1]
        idx = np.arange(nt - k)
        D[idx, idx + k] = c
        D[idx + k, idx] = -c
    D[:depth, :] = 0.0
    D[nt - depth:, :] = 0.0
    for r in range(depth):
        w = tab["widths"][r]
        row = np.asarray(tab["rows"][r], dtype=np.float64)
        D[r, :w] = row
        D[nt - 1 - r, nt - w:] = -row[::-1]
    return D.astype(np.float32)


def compute_time_derivative(Z: torch.Tensor, dt: float, fd_type: str = "sbp12") -> torch.Tensor:
    """dZ/dt = (1/dt) * D_fd @ Z in fp32.  Reference src/lasdi/latent_dynamics/sindy.py:89-102
    (`1. / Dt * torch.sparse.mm(self.fd_oper, Z)`); the operator is applied as a sparse COO
    product exactly as the reference does so that the fp32 accumulation order is the same."""
    D = torch.from_numpy(fd_operator_dense(Z.shape[0], fd_type)).to_sparse_coo()
    return 1.0 / dt * torch.sparse.mm(D, Z)


# ----------------------------------------------------------------------------------------
# sindy.py: calibrate / simulate
# ----------------------------------------------------------------------------------------

def calibrate_one(Z: torch.Tensor, dt: float, fd_type: str = "sbp12", coef_norm_order=1):
    """One training series.  Reference sindy.py:66-87: dZdt by the SBP stencil, library
    [1 | Z], `torch.linalg.lstsq` (LAPACK gelsy on CPU), MSE loss and coefficient norm.
    Returns (coefs flat fp32 numpy [(n+1)*n], loss_sindy float, loss_coef float)."""
    assert Z.dim() == 2
    Z = Z.to(torch.float32)
    dZdt = compute_time_derivative(Z, dt, fd_type)
    nt = Z.shape[0]
    Zi = torch.cat([torch.ones(nt, 1), Z], dim=1)
    coefs = torch.linalg.lstsq(Zi, dZdt).solution
    loss_sindy = torch.nn.functional.mse_loss(Zi @ coefs, dZdt)
    loss_coef = torch.norm(coefs, coef_norm_order)
    return coefs.detach().flatten().numpy(), float(loss_sindy), float(loss_coef)


def calibrate(Z: torch.Tensor, dt: float, fd_type: str = "sbp12", coef_norm_order=1):
    """Batch over training series; the two losses are SUMMED over the batch.
    Reference sindy.py:41-64.  Returns (coefs fp64 [n_train, ncoef], loss_sindy, loss_coef)."""
    if Z.dim() == 2:
        return calibrate_one(Z, dt, fd_type, coef_norm_order)
    n_train = Z.shape[0]
    ncoef = Z.shape[2] * (Z.shape[2] + 1)
    coefs = np.zeros([n_train, ncoef])
    ls, lc = 0.0, 0.0
    for i in range(n_train):
        c, a, b = calibrate_one(Z[i], dt, fd_type, coef_norm_order)
        coefs[i] = c
        ls += a
        lc += b
    return coefs, ls, lc


def calibrate_grad(Z: torch.Tensor, dt: float, fd_type: str = "sbp12", coef_norm_order=1, w_sindy: float = 1.0,
                   w_coef: float = 1.0, dtype=torch.float32) -> np.ndarray:
    """d(w_sindy * loss_sindy + w_coef * loss_coef)/dZ by torch autograd through the reference's op sequence
    (sparse.mm -> cat -> linalg.lstsq -> MSELoss / norm; reference sindy.py:66-77), losses summed over the
    batch as the reference's training loop consumes them (gplasdi.py:186-190).  `dtype=torch.float64` gives the
    round-off-free gradient used to judge fp32 results."""
    Zb = Z if Z.dim() == 3 else Z.unsqueeze(0)
    Zl = Zb.detach().to(dtype).clone().requires_grad_(True)
    total = 0.0
    for i in range(Zl.shape[0]):
        nt = Zl.shape[1]
        D = torch.from_numpy(fd_operator_dense(nt, fd_type)).to(dtype).to_sparse_coo()
        dZdt = 1.0 / dt * torch.sparse.mm(D, Zl[i])
        Zi = torch.cat([torch.ones(nt, 1, dtype=dtype), Zl[i]], dim=1)
        coefs = torch.linalg.lstsq(Zi, dZdt).solution
        total = total + w_sindy * torch.nn.functional.mse_loss(dZdt, Zi @ coefs) + w_coef * torch.norm(coefs, coef_norm_order)
    total.backward()
    g = Zl.grad.numpy()
    return g if Z.dim() == 3 else g[0]


def simulate(coefs: np.ndarray, z0: np.ndarray, t_grid: np.ndarray) -> np.ndarray:
    """One latent trajectory.  Reference sindy.py:104-117: coefficient vector = row-major
    flatten of the (n+1, n) lstsq solution; row 0 is the constant term; LSODA via scipy
    `odeint` with default tolerances.  Returns fp64 [nt, n]."""
    n = z0.shape[0]
    c_i = np.asarray(coefs).reshape([n + 1, n]).T
    dzdt = lambda z, t: c_i[:, 1:] @ z + c_i[:, 0]
    return odeint(dzdt, z0, t_grid)


def library_terms(n: int, poly_order: int = 1, sine_term: bool = False) -> int:
    """|Theta| = 1 + n + lib_size(n, poly_order) + n sines.  Reference legacy/Vlasov1D1V/utils/utils_sindy.py:12-27."""
    from scipy.special import binom
    n_poly = sum(int(binom(n + k - 1, k)) for k in range(poly_order + 1)) - n - 1
    return 1 + n + n_poly + (n if sine_term else 0)


def simulate_library(coefs: np.ndarray, z0: np.ndarray, t_grid: np.ndarray, poly_order: int = 1,
                     sine_term: bool = False) -> np.ndarray:
    """One latent trajectory of dz/dt = Theta(z) R for the polynomial / sine candidate libraries.
    Reference legacy/Vlasov1D1V/utils/utils_sindy.py:91-170 (`simulate_sindy`, one iteration of its loop): `c_i = R.T`;
    Theta = [1, z, z_i z_j for i <= j (i slowest) if poly_order == 2, sin(z) if sine_term]; LSODA via scipy `odeint`
    with default tolerances.  `coefs` = R [|Theta|, n] (or its row-major flatten).  Returns fp64 [nt, n]."""
    n = z0.shape[0]
    c_i = np.asarray(coefs, dtype=np.float64).reshape([library_terms(n, poly_order, sine_term), n]).T
    n_q = n * (n + 1) // 2 if poly_order == 2 else 0

    def dzdt(z, t):
        parts = [z]
        if poly_order == 2:
            parts.append(np.array([z[i] * z[j] for i in range(n) for j in range(i, n)]))
        if sine_term:
            parts.append(np.sin(z))
        return c_i[:, 1:] @ np.concatenate(parts) + c_i[:, 0]

    assert c_i.shape[1] == 1 + n + n_q + (n if sine_term else 0)
    return odeint(dzdt, z0, t_grid)


def simulate_exact(coefs: np.ndarray, z0: np.ndarray, dt: float, nt: int) -> np.ndarray:
    """Independent fp64 cross-check of `simulate` for the affine ODE: exact propagator
    expm([[A, b], [0, 0]] * dt) applied step by step (not in the reference; used only to
    report the LSODA tolerance floor next to the GPU-vs-oracle gap)."""
    from scipy.linalg import expm
    n = z0.shape[0]
    c_i = np.asarray(coefs, dtype=np.float64).reshape([n + 1, n]).T
    M = np.zeros((n + 1, n + 1))
    M[:n, :n] = c_i[:, 1:]
    M[:n, n] = c_i[:, 0]
    E = expm(M * dt)
    out = np.zeros((nt, n))
    z = np.concatenate([np.asarray(z0, dtype=np.float64), [1.0]])
    for k in range(nt):
        out[k] = z[:n]
        z = E @ z
    return out


def rollout_ensemble(coef_samples: np.ndarray, Z0: np.ndarray, t_grid: np.ndarray) -> np.ndarray:
    """Zis[p, s] = simulate(coef_samples[p, s], Z0[p], t_grid).
    Reference src/lasdi/gplasdi.py:262-266 (same loop in sample_roms :43-48)."""
    P, S = coef_samples.shape[:2]
    n = Z0.shape[1]
    Zis = np.zeros([P, S, len(t_grid), n])
    for i in range(P):
        for j in range(S):
            Zis[i, j] = simulate(coef_samples[i, j], Z0[i], t_grid)
    return Zis


# ----------------------------------------------------------------------------------------
# latent_space.py: MLP decoder
# ----------------------------------------------------------------------------------------

def _activation(name: str):
    """torch-default activations, reference src/lasdi/latent_space.py:5-28 (`act_dict`)."""
    table = {
        "sigmoid": torch.sigmoid,
        "tanh": torch.tanh,
        "softplus": torch.nn.functional.softplus,
        "ReLU": torch.relu,
        "ELU": torch.nn.functional.elu,
        "SiLU": torch.nn.functional.silu,
        "GELU": torch.nn.functional.gelu,
        # the remaining parameter-free modules of act_dict, with the default arguments the reference constructs them with
        "hardshrink": torch.nn.Hardshrink(), "hardsigmoid": torch.nn.Hardsigmoid(), "hardtanh": torch.nn.Hardtanh(),
        "hardswish": torch.nn.Hardswish(), "leakyReLU": torch.nn.LeakyReLU(), "logsigmoid": torch.nn.LogSigmoid(),
        "ReLU6": torch.nn.ReLU6(), "SELU": torch.nn.SELU(), "CELU": torch.nn.CELU(), "mish": torch.nn.Mish(),
        "softshrink": torch.nn.Softshrink(), "tanhshrink": torch.nn.Tanhshrink(),
    }
    return table[name]


def decoder_forward(weights, biases, act: str, z: torch.Tensor) -> torch.Tensor:
    """fp32 MLP: act after every layer but the last.  Reference latent_space.py:157-164
    (`MultiLayerPerceptron.forward`); weights are `nn.Linear` layout W[out, in], b[out]."""
    f = _activation(act)
    x = z
    with torch.no_grad():
        for i in range(len(weights) - 1):
            x = f(torch.nn.functional.linear(x, weights[i], biases[i]))
        x = torch.nn.functional.linear(x, weights[-1], biases[-1])
    return x


# ----------------------------------------------------------------------------------------
# gplasdi.py: get_fom_max_std
# ----------------------------------------------------------------------------------------

def encode_initial_conditions(weights, biases, act: str, U0: np.ndarray) -> np.ndarray:
    """Z0[i] = encoder(u0_i), one parameter at a time with a [1, 1, D] input, as the loop of
    `initial_condition_latent` does (reference src/lasdi/latent_space.py:49-61; encoder forward :157-164).
    U0 [P, D] -> fp32 [P, n_z]."""
    a = _activation(act)
    Ws = [torch.as_tensor(np.asarray(w), dtype=torch.float32) for w in weights]
    bs = [torch.as_tensor(np.asarray(b), dtype=torch.float32) for b in biases]
    out = []
    with torch.no_grad():
        for u in np.asarray(U0):
            x = torch.Tensor(np.asarray(u).reshape(1, 1, -1))
            for l in range(len(Ws) - 1):
                x = a(torch.nn.functional.linear(x, Ws[l], bs[l]))
            x = torch.nn.functional.linear(x, Ws[-1], bs[-1])
            out.append(x[0, 0, :].numpy())
    return np.stack(out)


def fom_max_std(weights, biases, act: str, Zis: np.ndarray, t_chunk: int | None = None):
    """Per-parameter max over (t, dof) of the population std over samples, and the
    first-strict-maximum index.  Reference src/lasdi/gplasdi.py:52-74: fp64 -> fp32 cast of
    Z (`torch.Tensor(Zi)`), decode, `numpy.std(0)` (ddof=0, fp32), `.max()`, strict `>`
    update starting from max_std = 0.

    `t_chunk` decodes the time axis in slabs (exact: std is taken per (t, dof)); None =
    verbatim single shot.  Returns (m_index or None, max_std fp32 [P])."""
    weights = [torch.as_tensor(w, dtype=torch.float32) for w in weights]
    biases = [torch.as_tensor(b, dtype=torch.float32) for b in biases]
    P = Zis.shape[0]
    per_param = np.zeros(P, dtype=np.float32)
    max_std = 0
    m_index = None
    for m in range(P):
        Zi = Zis[m]
        T = Zi.shape[1]
        step = T if t_chunk is None else t_chunk
        best = np.float32(0)
        first = True
        for t0 in range(0, T, step):
            Z_m = torch.Tensor(Zi[:, t0:t0 + step])
            X = decoder_forward(weights, biases, act, Z_m).numpy()
            v = X.std(0).max()
            best = v if first else max(best, v)
            first = False
        per_param[m] = best
        if best > max_std:
            m_index = m
            max_std = best
    return m_index, per_param


# ----------------------------------------------------------------------------------------
# gp.py: coefficient draws
# ----------------------------------------------------------------------------------------

def sample_coefs_from_moments(pred_mean: np.ndarray, pred_std: np.ndarray, n_samples: int) -> np.ndarray:
    """coef[s, k] ~ N(mean_k, std_k) from numpy's GLOBAL legacy RNG in (s, k) order.
    Reference src/lasdi/gp.py:76-78 (scalar double loop).  One vectorised call consumes the
    MT19937 stream identically (verified bit-exact in tests/test_oracle_golden.py)."""
    return np.random.normal(pred_mean[None, :], pred_std[None, :], size=(n_samples, pred_mean.shape[0]))


def greedy_step(weights, biases, act, coef_samples, Z0, t_grid, t_chunk=None):
    """rollout -> decode -> std -> max -> argmax, reference gplasdi.py:262-268."""
    Zis = rollout_ensemble(coef_samples, Z0, t_grid)
    idx, per_param = fom_max_std(weights, biases, act, Zis, t_chunk)
    return idx, per_param, Zis
