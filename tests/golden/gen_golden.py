#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (imports /root/reference/src/lasdi).

Run in the build container only:  python tests/golden/gen_golden.py [--full]
The fixtures pin the CPU oracle (oracle/reference_path.py) and, through it, the CUDA path.
Inputs come from gplasdi_b200.synth (seeded); only reference OUTPUTS (and small inputs) are stored.

Reference entry points exercised (file:line):
  lasdi.fd.FDdict[*].getOperators                     src/lasdi/fd.py:14-44
  lasdi.latent_dynamics.sindy.SINDy.calibrate/simulate src/lasdi/latent_dynamics/sindy.py:41-117
  lasdi.latent_space.MultiLayerPerceptron.forward     src/lasdi/latent_space.py:135-171
  lasdi.gplasdi.get_fom_max_std                       src/lasdi/gplasdi.py:52-74
  lasdi.gp.fit_gps / eval_gp / sample_coefs           src/lasdi/gp.py:5-80
"""
import os, sys, argparse, time
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np
np.Inf = np.inf  # numpy-2 shim for reference gplasdi.py:141
import torch

from lasdi.fd import FDdict
from lasdi.latent_dynamics.sindy import SINDy
from lasdi.latent_space import MultiLayerPerceptron
from lasdi.gplasdi import get_fom_max_std
from lasdi.gp import fit_gps, eval_gp, sample_coefs

from gplasdi_b200 import synth


class _AE:  # duck-typed autoencoder: get_fom_max_std only touches .decoder
    def __init__(self, widths, act, Ws, bs):
        self.decoder = MultiLayerPerceptron(list(widths), act, reshape_index=-1, reshape_shape=[widths[-1]])
        sd = {}
        for i, (W, b) in enumerate(zip(Ws, bs)):
            sd[f"fcs.{i}.weight"] = torch.from_numpy(W)
            sd[f"fcs.{i}.bias"] = torch.from_numpy(b)
        self.decoder.load_state_dict(sd)


def gen_fd():
    out = {}
    for name in ["sbp12", "sbp24", "sbp36", "sbp48"]:
        for nt in [17, 33]:
            D, norm, _ = FDdict[name].getOperators(nt)
            out[f"{name}_{nt}"] = D.to_dense().numpy()
        D, _, _ = FDdict[name].getOperators(1001)
        Dd = D.to_dense().numpy()
        out[f"{name}_1001_head"] = Dd[:12, :24]
        out[f"{name}_1001_tail"] = Dd[-12:, -24:]
        out[f"{name}_1001_mid"] = Dd[500, 490:511]
        out[f"{name}_1001_nnz"] = np.array([D._nnz()])
    np.savez_compressed(os.path.join(HERE, "fd_operators.npz"), **out)


def gen_calibrate():
    out = {}
    for name in ["sbp12", "sbp24", "sbp36", "sbp48"]:
        for (B, T, order) in [(3, 65, 1), (2, 1001, "fro")]:
            dt = 1.0 / (T - 1)
            Z = synth.make_latent_series(B, T, 5, dt, seed=7)
            ld = SINDy(5, T, {"sindy": {"fd_type": name, "coef_norm_order": order}})
            coefs, ls, lc = ld.calibrate(torch.from_numpy(Z), dt, compute_loss=True, numpy=True)
            dz = ld.compute_time_derivative(torch.from_numpy(Z[0]), dt).numpy()
            key = f"{name}_B{B}_T{T}"
            out[key + "_coefs"] = coefs
            out[key + "_loss_sindy"] = np.array([float(ls)])
            out[key + "_loss_coef"] = np.array([float(lc)])
            out[key + "_dzdt0"] = dz
    np.savez_compressed(os.path.join(HERE, "calibrate.npz"), **out)


GRAD_CASES = [("sbp12", 3, 65, 1), ("sbp24", 2, 129, "fro"), ("sbp36", 2, 65, 1), ("sbp48", 2, 1001, "fro"),
              ("sbp12", 2, 1001, 1)]
GRAD_W = (0.7, 1.0e-3)          # weights of (loss_sindy, loss_coef): both terms contribute at the same order


def gen_calibrate_grad():
    """Z.grad after backpropagating the reference's own calibrate() losses (what train() does, gplasdi.py:186-197)."""
    out = {}
    for (name, B, T, order) in GRAD_CASES:
        dt = 1.0 / (T - 1)
        Z = torch.from_numpy(synth.make_noisy_series(B, T, 5, dt, seed=11)).requires_grad_(True)
        ld = SINDy(5, T, {"sindy": {"fd_type": name, "coef_norm_order": order}})
        coefs, ls, lc = ld.calibrate(Z, dt, compute_loss=True, numpy=True)
        (GRAD_W[0] * ls + GRAD_W[1] * lc).backward()
        key = f"{name}_B{B}_T{T}"
        out[key + "_grad"] = Z.grad.numpy().copy()
        out[key + "_loss_sindy"] = np.array([float(ls)])
        out[key + "_loss_coef"] = np.array([float(lc)])
    np.savez_compressed(os.path.join(HERE, "calibrate_grad.npz"), **out)


def rankdef_series(kind: str, T: int = 129) -> np.ndarray:
    """Latent series whose library matrix [1 | Z] is rank deficient the way an untrained / saturated encoder makes it."""
    dt = 1.0 / (T - 1)
    Z = synth.make_noisy_series(1, T, 5, dt, seed=21)[0].copy()
    if kind == "const_col":
        Z[:, 3] = 0.75                     # a constant latent column: collinear with the library's 1
    elif kind == "collinear":
        Z[:, 4] = 2.0 * Z[:, 1]            # two proportional latent columns
    elif kind == "two_lost":
        Z[:, 3] = -0.5; Z[:, 2] = Z[:, 0] - Z[:, 1]
    elif kind == "near":
        rs = np.random.RandomState(5)
        Z[:, 4] = Z[:, 1] + 3.0e-6 * rs.normal(size=T).astype(np.float32)      # cond ~ 1e6 >> 1/rcond: truncated by gelsy
    return np.ascontiguousarray(Z[None], dtype=np.float32)


RANKDEF_KINDS = ["const_col", "collinear", "two_lost", "near"]


def gen_calibrate_rankdef():
    """The reference's own calibrate() (torch.linalg.lstsq = LAPACK sgelsy: rank-truncated minimum-norm answer,
    sindy.py:72) and its autograd gradient on rank-deficient series.

    torch's CPU lstsq is NOT reproducible on these inputs: repeated calls on the same tensor return different
    solutions (three different losses in three calls were observed here; LAPACK's xGEQP3 treats non-zero entries of its
    JPVT argument as columns to pin to the front, and the answers look like that array is handed over uninitialised).
    The rule LAPACK documents -- JPVT zeroed, free pivoting -- is what scipy's sgelsy wrapper runs, deterministically.
    The fixture therefore stores the reference run (calibrate + backward, repeated up to 200 times) whose coefficients
    agree with scipy's sgelsy on the same fp32 matrices, together with the losses of the other answers that were seen."""
    import scipy.linalg as sl
    out = {}
    for kind in RANKDEF_KINDS:
        Zn = rankdef_series(kind)
        T = Zn.shape[1]; dt = 1.0 / (T - 1)
        ld = SINDy(5, T, {"sindy": {"fd_type": "sbp12", "coef_norm_order": 1}})
        dz = ld.compute_time_derivative(torch.from_numpy(Zn[0]), dt).numpy()
        X = np.concatenate([np.ones((T, 1), dtype=np.float32), Zn[0]], axis=1)
        x_l, _, rank_l, _ = sl.lstsq(X, dz, cond=float(np.finfo(np.float32).eps) * max(X.shape), lapack_driver="gelsy")
        seen = []
        for attempt in range(200):
            Z = torch.from_numpy(Zn).requires_grad_(True)
            coefs, ls, lc = ld.calibrate(Z, dt, compute_loss=True, numpy=True)
            seen.append(float(ls))
            if np.max(np.abs(coefs.ravel() - x_l.ravel())) <= 1e-4 * np.max(np.abs(x_l)):
                (GRAD_W[0] * ls + GRAD_W[1] * lc).backward()
                break
        else:
            raise RuntimeError(f"{kind}: no reference run agreed with sgelsy in 200 attempts; losses seen {sorted(set(seen))}")
        out[kind + "_Z"] = Zn
        out[kind + "_coefs"] = coefs
        out[kind + "_coefs_sgelsy"] = x_l.ravel(); out[kind + "_rank_sgelsy"] = np.array([rank_l])
        out[kind + "_loss_sindy"] = np.array([float(ls)]); out[kind + "_loss_coef"] = np.array([float(lc)])
        out[kind + "_grad"] = Z.grad.numpy().copy()
        out[kind + "_losses_seen"] = np.array(sorted(set(seen)))
        out[kind + "_sv"] = np.linalg.svd(X.astype(np.float64), compute_uv=False)
        print(kind, "sgelsy rank", rank_l, "accepted after", attempt + 1, "runs; losses seen", sorted(set(seen)),
              "grad finite", bool(np.isfinite(Z.grad.numpy()).all()))
    np.savez_compressed(os.path.join(HERE, "calibrate_rankdef.npz"), **out)


class _Phys:     # what initial_condition_latent reads (reference src/lasdi/physics/__init__.py:4-79)
    def __init__(self, nx):
        self.qgrid_size = [nx]
        self.x_grid = np.linspace(-3.0, 3.0, nx)

    def initial_condition(self, param):
        a, w = param
        return a * np.exp(-self.x_grid ** 2 / 2 / w / w)


ENC_CASES = [(301, [100], "sigmoid", 3), (257, [64, 32], "softplus", 4), (1001, [100], "tanh", 5)]


def gen_encoder():
    """Z0 of the reference's initial_condition_latent (its per-parameter loop) with the reference's own Autoencoder,
    initialised under torch.manual_seed; the weights are stored too (the CUDA tests load them verbatim)."""
    from lasdi.latent_space import Autoencoder, initial_condition_latent
    out = {}
    a = np.linspace(0.7, 0.9, 4); w = np.linspace(0.9, 1.1, 3)
    A, W = np.meshgrid(a, w, indexing="ij")
    grid = np.stack([A.ravel(), W.ravel()], axis=1)
    out["param_grid"] = grid
    for (nx, hidden, act, seed) in ENC_CASES:
        torch.manual_seed(seed)
        ae = Autoencoder(_Phys(nx), {"hidden_units": hidden, "latent_dimension": 5, "activation": act})
        Z0 = np.stack(initial_condition_latent(grid, _Phys(nx), ae))
        key = f"nx{nx}_{act}"
        out[key + "_Z0"] = Z0
        for k, v in ae.state_dict().items():
            if k.startswith("encoder."):
                out[key + "_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "encoder.npz"), **out)


def gen_train():
    """Loss histories / best_coefs of the reference's own BayesianGLaSDI.train() (CPU) on the tiny problem."""
    import tempfile
    from lasdi.latent_space import Autoencoder
    from lasdi.gplasdi import BayesianGLaSDI

    nt, nx, train_space, X, cfg = synth.make_train_problem()

    class Phys:
        qgrid_size = [nx]; dt = 1.0 / (nt - 1); t_grid = np.linspace(0.0, 1.0, nt)

    class PS:
        def n_train(self): return train_space.shape[0]
    PS.train_space = train_space
    torch.manual_seed(21)
    ae = Autoencoder(Phys(), {"hidden_units": [32], "latent_dimension": 5, "activation": "sigmoid"})
    out = {"init_" + k: v.numpy().copy() for k, v in ae.state_dict().items()}
    ld = SINDy(5, nt, {"sindy": {"fd_type": "sbp12", "coef_norm_order": "fro"}})
    # parameter gradients of the first iteration, formed exactly as the loop body of train() forms the loss
    # (gplasdi.py:179-190): the well-conditioned half of the comparison (Adam's first step is lr * sign(grad), which
    # turns round-off in near-zero gradient entries into O(lr) differences of the trajectory)
    Xt = torch.from_numpy(X)
    Z = ae.encoder(Xt)
    loss_ae = torch.nn.MSELoss()(Xt, ae.decoder(Z))
    _, l_ld, l_coef = ld.calibrate(Z, Phys.dt, compute_loss=True, numpy=True)
    (loss_ae + cfg["ld_weight"] * l_ld / 4 + cfg["coef_weight"] * l_coef / 4).backward()
    for k, v in ae.named_parameters():
        out["grad_" + k] = v.grad.numpy().copy()
    ae.zero_grad()
    tmp = tempfile.mkdtemp()
    tr = BayesianGLaSDI(Phys(), ae, ld, PS(), dict(cfg, path_checkpoint=tmp + "/ckpt", path_results=tmp + "/res"))
    tr.X_train = torch.from_numpy(X)
    tr.train()
    out["training_loss"] = np.array(tr.training_loss); out["ae_loss"] = np.array(tr.ae_loss)
    out["ld_loss"] = np.array(tr.ld_loss); out["coef_loss"] = np.array(tr.coef_loss)
    out["best_coefs"] = np.asarray(tr.best_coefs); out["best_loss"] = np.array([tr.best_loss])
    for k, v in ae.state_dict().items():
        out["final_" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **out)


def _rollout_ref(ld, coefs, z0, t_grid):
    P, S = coefs.shape[:2]
    Zis = np.zeros([P, S, len(t_grid), z0.shape[1]])
    for i in range(P):
        for j in range(S):
            Zis[i, j] = ld.simulate(coefs[i, j], z0[i], t_grid)
    return Zis


def gen_greedy(names, fname, keep_Z=True, p_sel=None, f64=False):
    out = {}
    for nm in names:
        wl = synth.WORKLOADS[nm]
        Ws, bs = synth.make_decoder(wl.widths, wl.seed)
        sl = None if p_sel is None else p_sel
        inp = synth.make_inputs(wl, sl)
        ld = SINDy(wl.n_z, wl.T, {"sindy": {"fd_type": "sbp12"}})
        t0 = time.perf_counter()
        Zis = _rollout_ref(ld, inp["coefs"], inp["z0"], inp["t_grid"])
        t1 = time.perf_counter()
        ae = _AE(wl.widths, wl.act, Ws, bs)
        m_index = get_fom_max_std(ae, Zis)
        # per-parameter values, computed exactly as the loop body of get_fom_max_std does
        per = np.zeros(Zis.shape[0], dtype=np.float32)
        per64 = np.zeros(Zis.shape[0], dtype=np.float64)
        import copy
        dec64 = copy.deepcopy(ae.decoder).double()
        for m, Zi in enumerate(Zis):
            X = ae.decoder(torch.Tensor(Zi)).detach().numpy()
            per[m] = X.std(0).max()
            if f64:
                # the same function in fp64 arithmetic on the SAME fp32-cast trajectories: how far the reference's own fp32
                # decode + fp32 two-pass std is from the exact value (its noise floor; a tolerance term of the GPU tests)
                per64[m] = dec64(torch.Tensor(Zi).double()).detach().numpy().std(0).max()
        t2 = time.perf_counter()
        out[nm + "_m_index"] = np.array([m_index])
        out[nm + "_max_std"] = per
        out[nm + "_index"] = inp["index"]
        if f64:
            out[nm + "_max_std_f64"] = per64
        if keep_Z:
            out[nm + "_Zis"] = Zis
        print(f"{nm}: P={Zis.shape[0]} index={m_index} rollout {t1-t0:.1f}s decode+std {t2-t1:.1f}s "
              f"top2 margin={(np.sort(per)[-1]-np.sort(per)[-2])/np.sort(per)[-1]:.3%}")
    np.savez_compressed(os.path.join(HERE, fname), **out)


def gen_gp():
    """fit_gps/eval_gp/sample_coefs on the 4 corners of the shipped example."""
    rs = np.random.RandomState(11)
    X = np.array([[0.7, 0.9], [0.7, 1.1], [0.9, 0.9], [0.9, 1.1]])
    Y = rs.normal(size=(4, 30))
    gps = fit_gps(X, Y)
    test = np.array([[0.74, 0.98], [0.8, 1.0], [0.9, 1.1]])
    # single-point predicts, exactly as sample_coefs evaluates them (gp.py:73)
    ms = [eval_gp(gps, test[i]) for i in range(len(test))]
    mean = np.concatenate([m for m, _ in ms]); std = np.concatenate([s for _, s in ms])
    np.random.seed(5)
    samples = [sample_coefs(gps, test[i], 6) for i in range(len(test))]
    np.savez_compressed(os.path.join(HERE, "gp_sampling.npz"), X=X, Y=Y, test=test, mean=mean, std=std,
                        samples=np.stack(samples), seed=np.array([5]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run config #1 (burgers1d_shipped, ~1 min)")
    ap.add_argument("--only", default=None, help="regenerate one fixture family (e.g. calibrate_grad)")
    args = ap.parse_args()
    torch.manual_seed(0)
    if args.only == "shipped":
        gen_greedy(["burgers1d_shipped"], "greedy_burgers1d_shipped.npz", keep_Z=False, f64=True)
        print("done")
        sys.exit(0)
    if args.only:
        globals()["gen_" + args.only]()
        print("done")
        sys.exit(0)
    gen_fd()
    gen_calibrate()
    gen_calibrate_grad()
    gen_calibrate_rankdef()
    gen_encoder()
    gen_train()
    gen_gp()
    gen_greedy(["tiny", "small", "small_s200", "small_softplus"], "greedy_small.npz", keep_Z=True)
    if args.full:
        gen_greedy(["burgers1d_shipped"], "greedy_burgers1d_shipped.npz", keep_Z=False, f64=True)
    print("done")
