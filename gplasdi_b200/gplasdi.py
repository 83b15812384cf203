"""Greedy-sampling trainer front end on the B200 path; drop-in for the sampler functions of the
reference trainer module (reference src/lasdi/gplasdi.py:9-74, :241-274).

`BayesianGLaSDI.get_new_sample_point` keeps its signature and return value (np.ndarray
[1, n_param]) but hands the whole ensemble to ONE fused call (rollout -> decode -> variance ->
max -> argmax) instead of the reference's python loops; `get_fom_max_std`, `sample_roms` and
`average_rom` are the literal drop-ins for callers that pass their own arrays.

`train` is the reference's training leg with the SINDy calibration and its losses on
the CUDA kernels (analytic backward) -- checkpoints and `best_coefs` are interchangeable with the reference's.
"""
from __future__ import annotations

import numpy as np
import torch

from .gp import eval_gp, eval_gp_device, fit_gps, sample_coefs, sample_coefs_all, sample_coefs_all_device
from .latent_space import initial_condition_latent, mse_loss
from .engine import default_engine
from . import sharding


def _engine_for(autoencoder):
    dec = autoencoder.decoder
    if hasattr(dec, "_engine_with_weights"):
        dec._require_cuda_decoder()
        return dec._engine_with_weights()
    # foreign (e.g. the reference's own) MultiLayerPerceptron: read its nn.Linear stack
    eng = default_engine()
    eng.set_decoder([fc.weight for fc in dec.fcs], [fc.bias for fc in dec.fcs], dec.act_type)
    eng.decoder_owner = dec
    return eng


def average_rom(autoencoder, physics, latent_dynamics, gp_dictionary, param_grid):
    """Mean-coefficient rollout per test parameter, [n_test, nt, n_z] (reference gplasdi.py:9-23)."""
    param_grid = np.asarray(param_grid)
    if param_grid.ndim == 1:
        param_grid = param_grid.reshape(1, -1)
    Z0 = np.stack(initial_condition_latent(param_grid, physics, autoencoder))
    pred_mean, _ = eval_gp_device(gp_dictionary, param_grid, latent_dynamics.engine)
    return latent_dynamics.sample(pred_mean.cpu().numpy(), Z0, physics.t_grid)


def sample_roms(autoencoder, physics, latent_dynamics, gp_dictionary, param_grid, n_samples):
    """[n_test, n_samples, nt, n_z] ROM trajectories (reference gplasdi.py:25-50), one rollout launch."""
    param_grid = np.asarray(param_grid)
    if param_grid.ndim == 1:
        param_grid = param_grid.reshape(1, -1)
    Z0 = np.stack(initial_condition_latent(param_grid, physics, autoencoder))
    # posterior moments and the draws on the device, from numpy's global legacy stream in the order of the
    # reference's loop (parameter, then sample, then coefficient; gplasdi.py:43-44 -> gp.py:76-78)
    coef_samples = sample_coefs_all_device(gp_dictionary, param_grid, n_samples, latent_dynamics.engine)
    Z = latent_dynamics.rollout_ensemble(coef_samples, Z0, physics.t_grid)
    return Z.cpu().numpy()


def fom_max_std(autoencoder, Zis):
    """(m_index, max_std[P]) for caller-supplied trajectories Zis [P, S, nt, n_z]."""
    eng = _engine_for(autoencoder)
    Zt = torch.as_tensor(np.ascontiguousarray(Zis))
    ms, am, _ = eng.checked(lambda: eng.decode_max_std(Zt))
    return int(am.item()), ms.cpu().numpy()


def fom_mean_std(autoencoder, Zis):
    """(mean, std) fields [P, nt, *qgrid_size] of the decoded ensemble Zis [P, S, nt, n_z] -- what
    `plot_prediction` computes with `pred.mean(0)`, `pred.std(0)` (reference src/lasdi/postprocess.py:32-34),
    for every parameter at once and without materialising `pred`."""
    eng = _engine_for(autoencoder)
    Zt = torch.as_tensor(np.ascontiguousarray(Zis))
    mean, std = eng.decode_stat_fields(Zt)
    eng.check_numeric()
    q = list(getattr(autoencoder, "qgrid_size", [mean.shape[-1]]))
    P, T = mean.shape[:2]
    return mean.cpu().numpy().reshape(P, T, *q), std.cpu().numpy().reshape(P, T, *q)


def get_fom_max_std(autoencoder, Zis):
    """Index of the test parameter with the largest decoded sample std (reference gplasdi.py:52-74)."""
    m_index, _ = fom_max_std(autoencoder, Zis)
    if m_index < 0:
        # the reference leaves m_index unbound here and raises UnboundLocalError
        raise RuntimeError("no test parameter has a positive standard deviation")
    return m_index


class BayesianGLaSDI:
    """Sampler half of the reference trainer (reference gplasdi.py:91-294): same constructor
    arguments and config keys; `get_new_sample_point` is the B200 hot path."""

    def __init__(self, physics, autoencoder, latent_dynamics, param_space, config):
        self.autoencoder = autoencoder
        self.latent_dynamics = latent_dynamics
        self.physics = physics
        self.param_space = param_space
        self.n_samples = config["n_samples"]
        for k in ("lr", "n_iter", "max_iter", "max_greedy_iter", "ld_weight", "coef_weight"):
            setattr(self, k, config.get(k))
        self.path_checkpoint = config.get("path_checkpoint", "checkpoint")
        self.path_results = config.get("path_results", "results")
        self.device = config.get("device", "cuda")
        self.verbose = config.get("verbose", True)
        self.optimizer = torch.optim.Adam(autoencoder.parameters(), lr=self.lr)
        self.training_loss, self.ae_loss, self.ld_loss, self.coef_loss = [], [], [], []
        self.timer = _TimerStub()
        self.best_loss = np.inf
        self.best_coefs = None
        self.restart_iter = 0
        self.X_train = torch.Tensor([])
        self.X_test = torch.Tensor([])
        self.last_max_std = None

    def train(self):
        """One training leg (reference gplasdi.py:150-238): `n_iter` Adam steps on
        loss = MSE(X, decoder(encoder(X))) + ld_weight * loss_sindy / n_train + coef_weight * loss_coef / n_train,
        best-loss checkpoint + `best_coefs`.  The encoder/decoder passes that record an autograd graph are torch ops
        (cuBLAS) ONLY on the host; on the GPU the autoencoder's forward / backward and the reconstruction loss are this
        package's kernels too (glasdi_mlp_forward_train / glasdi_mlp_backward / glasdi_mse_loss_grad, one autograd node per
        network); the SINDy calibration and its two losses are this package's CUDA kernels with an analytic backward
        (`glasdi_sindy_fit` / `glasdi_sindy_fit_backward`), so the latent series never leaves the GPU -- the
        reference moves Z to the host every iteration (gplasdi.py:183) and differentiates through `lstsq` there."""
        import os
        assert self.X_train.size(0) > 0
        assert self.X_train.size(0) == self.param_space.n_train()
        dev = torch.device(self.device)
        ae = self.autoencoder.to(dev)
        X = self.X_train.to(dev)
        os.makedirs(self.path_checkpoint, exist_ok=True)
        os.makedirs(self.path_results, exist_ok=True)
        ckpt = os.path.join(self.path_checkpoint, "checkpoint.pt")
        n_train = self.param_space.n_train()
        ld = self.latent_dynamics
        _optimizer_to(self.optimizer, dev)
        self.training_loss, self.ae_loss, self.ld_loss, self.coef_loss = [], [], [], []      # per leg (reference :166-169)

        def save_checkpoint():
            torch.save({k: v.detach().cpu() for k, v in ae.state_dict().items()}, ckpt)

        coefs = None
        next_iter = min(self.restart_iter + self.n_iter, self.max_iter)
        for it in range(self.restart_iter, next_iter):
            self.optimizer.zero_grad()
            Z = ae.encoder(X)
            X_pred = ae.decoder(Z)
            loss_ae = mse_loss(X_pred, X)          # loss + gradient in one kernel pass (glasdi_mse_loss_grad)
            coefs, loss_ld, loss_coef = ld.calibrate(Z, self.physics.dt, compute_loss=True, numpy=True)
            loss = loss_ae + self.ld_weight * loss_ld / n_train + self.coef_weight * loss_coef / n_train
            # one device->host read per iteration for the four histories and the best-loss test
            lv, la, ll, lc = torch.stack([loss.detach(), loss_ae.detach(), loss_ld.detach(), loss_coef.detach()]).tolist()
            if not np.isfinite(lv):
                # never let a non-finite loss reach the optimiser: one step would poison every autoencoder weight
                raise FloatingPointError(f"non-finite training loss at iteration {it + 1} (AE {la}, LD {ll}, COEF {lc})")
            loss.backward()
            self.optimizer.step()
            self.training_loss.append(lv); self.ae_loss.append(la); self.ld_loss.append(ll); self.coef_loss.append(lc)
            if lv < self.best_loss:
                # as in the reference, the checkpoint is written AFTER the optimiser step of the best iteration
                save_checkpoint()
                self.best_coefs = coefs
                self.best_loss = lv
            if self.verbose:
                print("Iter: %05d/%d, Loss: %3.10f, Loss AE: %3.10f, Loss LD: %3.10f, Loss COEF: %3.10f, max|c|: %04.1f"
                      % (it + 1, self.max_iter, lv, la, ll, lc, np.abs(coefs).max()))
        self.restart_iter += self.n_iter
        if self.best_coefs is not None and self.best_coefs.shape[0] == n_train:
            self.autoencoder.load_state_dict(torch.load(ckpt))
        else:
            self.best_coefs = coefs
            save_checkpoint()

    def load_checkpoint(self):
        """Best-loss decoder weights written by the training loop (reference gplasdi.py:254)."""
        import os
        path = os.path.join(self.path_checkpoint, "checkpoint.pt")
        if os.path.exists(path):
            self.autoencoder.load_state_dict(torch.load(path))

    def get_new_sample_point(self, group=None):
        """Reference gplasdi.py:241-274.  With torch.distributed initialised the test grid is
        sharded across ranks; every rank returns the same parameter."""
        assert self.best_coefs is not None
        ps = self.param_space
        assert self.best_coefs.shape[0] == ps.n_train()
        n_test = ps.n_test()
        print("\n~~~~~~~ Finding New Point ~~~~~~~")
        ae = self.autoencoder
        self.load_checkpoint()

        import time
        import torch.distributed as dist
        from .latent_dynamics.sindy import _uniform_dt
        tm = {}
        t0 = time.perf_counter()
        Z0 = np.stack(initial_condition_latent(ps.test_space, self.physics, ae))
        tm["initial_condition_latent_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        gp_dictionnary = fit_gps(ps.train_space, self.best_coefs)
        tm["fit_gps_host_s"] = time.perf_counter() - t0             # scikit-learn hyper-parameter fit: host, out of scope
        eng = _engine_for(ae)
        # posterior moments and the (p, s, k)-ordered legacy-stream draws on the device: [n_test, n_samples, n_coef] fp64
        t0 = time.perf_counter()
        coef_samples = sample_coefs_all_device(gp_dictionnary, ps.test_space, self.n_samples, eng)
        torch.cuda.synchronize()
        tm["gp_posterior_and_draws_s"] = time.perf_counter() - t0

        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        bounds = sharding.shard_bounds(n_test, world)
        lo, hi = bounds[rank], bounds[rank + 1]

        dt = _uniform_dt(self.physics.t_grid)      # a non-uniform grid raises here instead of rolling out wrongly
        t0 = time.perf_counter()
        ms, key = sharding.sharded_greedy_step(eng, coef_samples[lo:hi], Z0[lo:hi], len(self.physics.t_grid), dt,
                                               lo, group)
        key_host = int(key.item())
        tm["greedy_step_and_allreduce_s"] = time.perf_counter() - t0
        self.last_timings = tm
        self.last_max_std = ms.cpu().numpy()
        _, m_index = sharding.unpack_maxloc(key_host)
        if m_index < 0:
            raise RuntimeError("no test parameter has a positive standard deviation")
        new_sample = ps.test_space[m_index, :].reshape(1, -1)
        print("New param: " + str(np.round(new_sample, 4)) + "\n")
        return new_sample

    def export(self):
        """Same keys as the reference trainer's dict (reference gplasdi.py:276-287), so that a restart file written here
        loads there and vice versa: optimiser state (Adam moments), timer dict and the four loss histories included."""
        return {"X_train": self.X_train, "X_test": self.X_test, "lr": self.lr, "n_iter": self.n_iter,
                "n_samples": self.n_samples, "best_coefs": self.best_coefs, "max_iter": self.max_iter,
                "max_greedy_iter": self.max_greedy_iter, "ld_weight": self.ld_weight, "coef_weight": self.coef_weight,
                "restart_iter": self.restart_iter, "timer": self.timer.export(),
                "optimizer": self.optimizer.state_dict(), "training_loss": self.training_loss, "ae_loss": self.ae_loss,
                "ld_loss": self.ld_loss, "coeff_loss": self.coef_loss}    # 'coeff_loss': the reference's spelling

    def load(self, dict_):
        """Reference gplasdi.py:289-300: restores the data, `best_coefs`, the restart counter, the timer and the optimiser
        state (moved to the training device); the loss histories when the dict carries them."""
        self.X_train = dict_["X_train"]
        self.X_test = dict_["X_test"]
        self.best_coefs = dict_["best_coefs"]
        self.restart_iter = dict_["restart_iter"]
        if "timer" in dict_:
            self.timer.load(dict_["timer"])
        if "optimizer" in dict_:
            self.optimizer.load_state_dict(dict_["optimizer"])
            if self.device != "cpu" and torch.cuda.is_available():
                _optimizer_to(self.optimizer, torch.device(self.device))
        for k, attr in (("training_loss", "training_loss"), ("ae_loss", "ae_loss"), ("ld_loss", "ld_loss"),
                        ("coeff_loss", "coef_loss")):
            if k in dict_:
                setattr(self, attr, list(dict_[k]))


def _optimizer_to(optim, device):
    """Moves the optimiser's state tensors (Adam moments) to `device` (reference gplasdi.py:77-89)."""
    for state in optim.state.values():
        for k, v in state.items():
            if isinstance(v, torch.Tensor):
                state[k] = v.to(device)


class _TimerStub:
    """Carries the reference's timer dict through export()/load() (reference src/lasdi/timing.py: `export` returns
    {'names': {name: index}, 'calls': [...], 'times': [...]}); timing itself is out of scope."""

    def __init__(self):
        self.state = {"names": {}, "calls": [], "times": []}

    def export(self):
        return self.state

    def load(self, d):
        self.state = dict(d)
