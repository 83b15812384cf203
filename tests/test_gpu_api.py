"""Drop-in Python API on the GPU: the classes and functions keep the reference's names, argument
meaning, return types and error behaviour (reference src/lasdi/latent_dynamics/sindy.py,
src/lasdi/latent_space.py, src/lasdi/gplasdi.py), so these tests read like tests of the reference."""
import os

import numpy as np
import pytest
import torch

from gplasdi_b200 import synth
from gplasdi_b200.gplasdi import BayesianGLaSDI, get_fom_max_std, fom_max_std, fom_mean_std, sample_roms, average_rom
from gplasdi_b200.gp import fit_gps, eval_gp
from gplasdi_b200.latent_dynamics.sindy import SINDy
from gplasdi_b200.latent_space import Autoencoder, initial_condition_latent
from oracle import reference_path as orc

pytestmark = pytest.mark.gpu


class Physics:
    """Minimal physics object: what the sampler reads (reference src/lasdi/physics/__init__.py:4-79)."""

    def __init__(self, nt=65, nx=301, tmax=1.0):
        self.qgrid_size = [nx]
        self.nt = nt
        self.t_grid = np.linspace(0.0, tmax, nt)
        self.dt = tmax / (nt - 1)
        self.x_grid = np.linspace(-3.0, 3.0, nx)

    def initial_condition(self, param):
        a, w = param
        return a * np.exp(-self.x_grid ** 2 / 2 / w / w)


class ParamSpace:
    def __init__(self):
        a = np.linspace(0.7, 0.9, 4); w = np.linspace(0.9, 1.1, 3)
        A, W = np.meshgrid(a, w, indexing="ij")
        self.test_space = np.stack([A.ravel(), W.ravel()], axis=1)
        self.train_space = np.array([[0.7, 0.9], [0.7, 1.1], [0.9, 0.9], [0.9, 1.1]])

    def n_test(self):
        return self.test_space.shape[0]

    def n_train(self):
        return self.train_space.shape[0]


def _autoencoder(physics, hidden=(100,), act="sigmoid", seed=0):
    torch.manual_seed(seed)
    return Autoencoder(physics, {"hidden_units": list(hidden), "latent_dimension": 5, "activation": act})


def test_sindy_simulate_and_sample(golden_dir):
    g = np.load(os.path.join(golden_dir, "greedy_small.npz"))
    wl = synth.WORKLOADS["small"]
    inp = synth.make_inputs(wl)
    ld = SINDy(5, wl.T, {"sindy": {"fd_type": "sbp12"}})
    Z = ld.simulate(inp["coefs"][3, 4], inp["z0"][3], inp["t_grid"])
    ref = g["small_Zis"][3, 4]
    assert isinstance(Z, np.ndarray) and Z.dtype == np.float64 and Z.shape == (wl.T, 5)
    assert np.max(np.abs(Z - ref)) <= 1e-5 * np.max(np.abs(ref))
    Zs = ld.sample(inp["coefs"][:, 0], inp["z0"], inp["t_grid"])
    assert Zs.shape == (wl.P, wl.T, 5)
    assert np.max(np.abs(Zs - g["small_Zis"][:, 0])) <= 1e-5 * np.max(np.abs(g["small_Zis"]))


def test_sindy_calibrate_return_contract(golden_dir):
    g = np.load(os.path.join(golden_dir, "calibrate.npz"))
    T = 65; dt = 1.0 / (T - 1)
    Z = torch.from_numpy(synth.make_latent_series(3, T, 5, dt, seed=7))
    ld = SINDy(5, T, {"sindy": {"fd_type": "sbp24", "coef_norm_order": 1}})
    coefs, ls, lc = ld.calibrate(Z, dt, compute_loss=True, numpy=True)
    assert isinstance(coefs, np.ndarray) and coefs.dtype == np.float64 and coefs.shape == (3, 30)
    ref = g["sbp24_B3_T65_coefs"]
    assert np.max(np.abs(coefs - ref)) <= 1e-5 * np.max(np.abs(ref))
    np.testing.assert_allclose(ls.item(), g["sbp24_B3_T65_loss_sindy"][0], rtol=1e-4, atol=5e-11)
    np.testing.assert_allclose(lc.item(), g["sbp24_B3_T65_loss_coef"][0], rtol=1e-5)
    c1 = ld.calibrate(Z[1], dt, compute_loss=False, numpy=False)             # one series, torch out
    assert isinstance(c1, torch.Tensor) and c1.dtype == torch.float32 and c1.shape == (30,)
    np.testing.assert_array_equal(c1.numpy().astype(np.float64), coefs[1])
    c3 = ld.calibrate(Z, dt, compute_loss=False)
    assert isinstance(c3, torch.Tensor) and c3.shape == (3, 30)
    dz = ld.compute_time_derivative(Z[0], dt)
    assert dz.shape == (T, 5) and dz.device.type == "cpu"
    # fitted coefficients reproduce the series they were fitted to when integrated
    Zsim = ld.simulate(coefs[0], Z[0, 0].numpy(), np.arange(T) * dt)
    assert np.max(np.abs(Zsim - Z[0].numpy())) < 5e-2 * np.max(np.abs(Z[0].numpy()))


def test_sindy_calibrate_is_differentiable_like_the_reference(golden_dir):
    """The reference's training loop backpropagates calibrate()'s losses into the encoder (gplasdi.py:186-197):
    same call, same Z.grad (fixture = the reference's own autograd), through an encoder-like upstream op too."""
    g = np.load(os.path.join(golden_dir, "calibrate_grad.npz"))
    T = 65; dt = 1.0 / (T - 1)
    Z0 = torch.from_numpy(synth.make_noisy_series(3, T, 5, dt, seed=11))
    ld = SINDy(5, T, {"sindy": {"fd_type": "sbp12", "coef_norm_order": 1}})
    for dev in ("cpu", "cuda"):
        Z = Z0.detach().clone().to(dev).requires_grad_(True)
        coefs, ls, lc = ld.calibrate(Z, dt, compute_loss=True, numpy=True)
        assert isinstance(coefs, np.ndarray) and coefs.dtype == np.float64          # detached, as in the reference
        assert ls.requires_grad and lc.requires_grad and ls.device == Z.device
        (0.7 * ls + 1.0e-3 * lc).backward()
        ref = g["sbp12_B3_T65_grad"]
        assert np.max(np.abs(Z.grad.cpu().numpy() - ref)) <= 1e-5 * np.max(np.abs(ref))
    # chain rule through an upstream op, one series, Frobenius norm
    ld2 = SINDy(5, T, {"sindy": {"fd_type": "sbp24", "coef_norm_order": "fro"}})
    w = torch.linspace(0.5, 1.5, 5).requires_grad_(True)
    out = ld2.calibrate(Z0[1] * w, dt, compute_loss=True)
    (out[1] + 0.01 * out[2]).backward()
    from oracle import reference_path as orc
    gz = orc.calibrate_grad(Z0[1] * w.detach(), dt, "sbp24", "fro", 1.0, 0.01, dtype=torch.float64)
    gw = (gz * Z0[1].numpy()).sum(0)
    assert np.max(np.abs(w.grad.numpy() - gw)) <= 1e-4 * np.max(np.abs(gw))
    # no graph requested -> plain (non-differentiable) losses, no autograd node
    with torch.no_grad():
        _, ls, lc = ld.calibrate(Z0.requires_grad_(True), dt)
    assert not ls.requires_grad


@pytest.mark.parametrize("act", ["sigmoid", "tanh", "softplus", "SiLU", "GELU", "ELU", "mish", "leakyReLU"])
@pytest.mark.parametrize("sizes,rows", [([37, 20, 5], 70), ([5, 100, 1001], 200), ([64, 33, 17, 4], 129)])
def test_training_mode_mlp_matches_torch_autograd(act, sizes, rows):
    """glasdi_mlp_forward_train / glasdi_mlp_backward behind one autograd node (MultiLayerPerceptron.forward with grad on
    the GPU) against the per-layer graph torch records for the same module (reference latent_space.py:157-164): output,
    input gradient and every parameter gradient.  Tolerance 2e-5 of the largest entry (fp32, different summation orders)."""
    from gplasdi_b200.latent_space import MultiLayerPerceptron
    torch.manual_seed(3)
    mlp = MultiLayerPerceptron(sizes, act).cuda()
    x = torch.randn(rows, sizes[0], device="cuda", requires_grad=True)
    gy = torch.randn(rows, sizes[-1], device="cuda")
    y = mlp(x)
    node = y.grad_fn
    while node is not None and not type(node).__name__.startswith("_MlpTrain"):
        node = node.next_functions[0][0] if node.next_functions else None
    assert node is not None                                          # the CUDA node, not torch's per-layer graph
    y.backward(gy)
    got = [x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    x.grad = None
    for p in mlp.parameters():
        p.grad = None
    h = x
    for i in range(len(mlp.fcs) - 1):
        h = mlp.act(mlp.fcs[i](h))
    yr = mlp.fcs[-1](h)
    yr.backward(gy)
    ref = [x.grad] + [p.grad for p in mlp.parameters()]
    assert float((y - yr).abs().max()) <= 2e-5 * float(yr.abs().max())
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-5 * max(float(b.abs().max()), 1e-30)


def test_mse_loss_kernel_matches_torch():
    from gplasdi_b200.latent_space import mse_loss
    torch.manual_seed(5)
    a = torch.randn(4, 1001, 257, device="cuda", requires_grad=True)
    b = torch.randn(4, 1001, 257, device="cuda")
    loss = mse_loss(a, b)
    (3.0 * loss).backward()
    g = a.grad.clone(); a.grad = None
    ref = torch.nn.functional.mse_loss(a.double(), b.double())
    (3.0 * ref).backward()
    assert abs(float(loss) - float(ref)) <= 1e-6 * float(ref)
    assert float((g - a.grad).abs().max()) <= 1e-6 * float(a.grad.abs().max())
    assert float(mse_loss(a.detach(), b)) == float(loss)               # bit-reproducible


def test_train_leg_matches_reference_trainer(golden_dir, tmp_path):
    """BayesianGLaSDI.train(): same loss histories, best coefficients and checkpoint as the reference's own train()
    (fixture, run on CPU by tests/golden/gen_golden.py::gen_train) from the same initial weights -- with the SINDy
    calibration forward/backward on the CUDA kernels and Z never leaving the device."""
    g = np.load(os.path.join(golden_dir, "train.npz"))
    nt, nx, train_space, X, cfg = synth.make_train_problem()

    class Phys:
        qgrid_size = [nx]; dt = 1.0 / (nt - 1); t_grid = np.linspace(0.0, 1.0, nt)

    class PS:
        def n_train(self): return train_space.shape[0]
    PS.train_space = train_space
    ae = Autoencoder(Phys(), {"hidden_units": [32], "latent_dimension": 5, "activation": "sigmoid"})
    ae.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init_")})
    ld = SINDy(5, nt, {"sindy": {"fd_type": "sbp12", "coef_norm_order": "fro"}})
    tr = BayesianGLaSDI(Phys(), ae, ld, PS(), dict(cfg, path_checkpoint=str(tmp_path / "ckpt"),
                                                   path_results=str(tmp_path / "res"), verbose=False))
    tr.X_train = torch.from_numpy(X)
    # (1) parameter gradients of the first iteration, formed as the loop body of train() forms the loss
    aed = ae.to("cuda")
    Xd = tr.X_train.to("cuda")
    Z = aed.encoder(Xd)
    loss_ae = torch.nn.MSELoss()(Xd, aed.decoder(Z))
    _, l_ld, l_coef = ld.calibrate(Z, Phys.dt, compute_loss=True, numpy=True)
    (loss_ae + cfg["ld_weight"] * l_ld / 4 + cfg["coef_weight"] * l_coef / 4).backward()
    # fp64 evaluation of the same graph (oracle ops on the CPU): at initialisation the latent series are nearly
    # collinear (cond([1|Z]) ~ 1e4), where the reference's OWN fp32 autograd through lstsq is 1-100 % off for the
    # encoder tensors.  The bar: at least as close to the fp64 gradient as the reference's fp32 result (fixture).
    from oracle import reference_path as orc
    ae64 = Autoencoder(Phys(), {"hidden_units": [32], "latent_dimension": 5, "activation": "sigmoid"}).double()
    ae64.load_state_dict({k[5:]: torch.from_numpy(g[k]).double() for k in g.files if k.startswith("init_")})
    X64 = tr.X_train.double()
    Z64 = ae64.encoder(X64)
    tot = torch.nn.MSELoss()(X64, ae64.decoder(Z64))
    D = torch.from_numpy(orc.fd_operator_dense(nt, "sbp12")).double()
    for i in range(4):
        dZ = 1.0 / Phys.dt * (D @ Z64[i])
        Zi = torch.cat([torch.ones(nt, 1, dtype=torch.float64), Z64[i]], dim=1)
        c = torch.linalg.lstsq(Zi, dZ).solution
        tot = tot + cfg["ld_weight"] / 4 * torch.nn.functional.mse_loss(dZ, Zi @ c) + cfg["coef_weight"] / 4 * torch.norm(c, "fro")
    tot.backward()
    truth = {k: v.grad.numpy() for k, v in ae64.named_parameters()}
    for k, v in aed.named_parameters():
        scale = np.max(np.abs(truth[k]))
        err_ref = np.max(np.abs(g["grad_" + k] - truth[k])) / scale
        err = np.max(np.abs(v.grad.cpu().numpy() - truth[k])) / scale
        print(f"{k}: this package {err:.2e}, reference fp32 {err_ref:.2e} (relative to the fp64 gradient)")
        assert err <= max(1e-4, 1.5 * err_ref), k
    aed.zero_grad()
    # (2) the training leg itself.  Adam's first steps are lr * sign(grad), and the encoder gradients of BOTH
    #     implementations carry the round-off measured above: the trajectories agree to what that allows (the
    #     coefficient norm and the tiny SINDy residual, 2e-5, feel it at the 1 % level after six steps)
    n0 = ld.engine.launch_count
    tr.train()
    # every kernel of an iteration is this package's: SINDy fit + its backward (2), encoder and decoder forward (2 GEMMs
    # each), the reconstruction loss (2), decoder backward (dW + its split-K reduction, db, dIn for both layers: 8), encoder
    # backward (the input needs no gradient: 7) -- and nothing else launches on the library's handle
    assert ld.engine.launch_count - n0 == 23 * cfg["n_iter"]
    np.testing.assert_allclose(tr.ld_loss[0], g["ld_loss"][0], rtol=1e-4)
    for name, rtol in (("training_loss", 2e-3), ("ae_loss", 2e-3), ("ld_loss", 3e-2), ("coef_loss", 2e-2)):
        np.testing.assert_allclose(getattr(tr, name), g[name], rtol=rtol, err_msg=name)
    assert tr.restart_iter == cfg["n_iter"] and abs(tr.best_loss - g["best_loss"][0]) <= 2e-3 * g["best_loss"][0]
    assert tr.best_coefs.shape == (4, 30) and tr.best_coefs.dtype == np.float64
    assert np.max(np.abs(tr.best_coefs - g["best_coefs"])) <= 3e-2 * np.max(np.abs(g["best_coefs"]))
    sd = torch.load(str(tmp_path / "ckpt" / "checkpoint.pt"))
    for k, v in sd.items():                                           # best-loss checkpoint, reloaded into the module
        ref = g["final_" + k]
        diff = np.abs(v.numpy() - ref)
        # an Adam step moves an entry by at most ~lr: entries whose (noise-level) gradient sign differs drift apart by
        # up to 2 lr per step, the bulk of the weights agrees closely
        assert np.max(diff) <= 2.2 * cfg["lr"] * cfg["n_iter"], k
        if k.startswith("decoder."):                                  # decoder gradients agree to 1e-7 (see above)
            assert np.mean(diff) <= 5e-4 * np.max(np.abs(ref)), k
        assert torch.equal(ae.state_dict()[k].cpu(), v)


def test_autoencoder_decoder_is_cuda_kernel_and_matches_torch():
    phys = Physics()
    ae = _autoencoder(phys)
    Z = torch.randn(4, 9, 5)
    with torch.no_grad():
        X = ae.decoder(Z)                                    # hot path: C-ABI glasdi_decode
    assert X.shape == (4, 9, 301) and X.device.type == "cpu"
    Ws = [fc.weight.detach() for fc in ae.decoder.fcs]; bs = [fc.bias.detach() for fc in ae.decoder.fcs]
    Xr = orc.decoder_forward(Ws, bs, "sigmoid", Z)
    assert torch.max(torch.abs(X - Xr)) <= 5e-6 * torch.max(torch.abs(Xr))
    n0 = ae.decoder._engine.launch_count
    with torch.no_grad():
        ae.decoder(Z)
    assert ae.decoder._engine.launch_count == n0 + 1         # one kernel, weights stay packed
    with torch.no_grad():
        ae.decoder.fcs[1].bias.add_(1.0)                     # weights changed -> repacked automatically
        X2 = ae.decoder(Z)
    assert torch.allclose(X2, X + 1.0, atol=1e-5)
    Xg = ae.decoder(Z.clone().requires_grad_(True))          # autograd requested: differentiable torch ops
    assert Xg.requires_grad


def test_initial_condition_latent_dropin(golden_dir):
    """Same call as the reference (latent_space.py:30-63): list of fp32 [n_z] vectors, from ONE batched CUDA encode."""
    from gplasdi_b200.latent_space import initial_condition_latent
    g = np.load(os.path.join(golden_dir, "encoder.npz"))
    phys = Physics(nx=301)
    ae = Autoencoder(phys, {"hidden_units": [100], "latent_dimension": 5, "activation": "sigmoid"})
    sd = ae.state_dict()
    for k in list(sd):
        if k.startswith("encoder."):
            sd[k] = torch.from_numpy(g["nx301_sigmoid_" + k])
    ae.load_state_dict(sd)
    from gplasdi_b200.engine import default_engine
    n0 = default_engine().launch_count
    Z0 = initial_condition_latent(g["param_grid"], phys, ae)
    assert isinstance(Z0, list) and len(Z0) == 12 and Z0[0].dtype == np.float32 and Z0[0].shape == (5,)
    ref = g["nx301_sigmoid_Z0"]
    assert np.max(np.abs(np.stack(Z0) - ref)) <= 1e-5 * np.max(np.abs(ref))
    assert ae.encoder._engine.launch_count - n0 == 2          # two layers, all 12 parameters at once
    Zg = ae.encoder(torch.zeros(2, 1, 301, requires_grad=True))      # autograd requested: torch ops
    assert Zg.requires_grad


def test_get_fom_max_std_dropin(golden_dir):
    g = np.load(os.path.join(golden_dir, "greedy_small.npz"))
    wl = synth.WORKLOADS["small"]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    phys = Physics(nt=wl.T, nx=wl.D)
    ae = _autoencoder(phys)
    sd = ae.state_dict()
    for i, (W, b) in enumerate(zip(Ws, bs)):
        sd[f"decoder.fcs.{i}.weight"] = torch.from_numpy(W); sd[f"decoder.fcs.{i}.bias"] = torch.from_numpy(b)
    ae.load_state_dict(sd)
    m = get_fom_max_std(ae, g["small_Zis"])
    assert isinstance(m, int) and m == int(g["small_m_index"][0])
    _, per = fom_max_std(ae, g["small_Zis"])
    assert np.max(np.abs(per - g["small_max_std"]) - 1e-5 * g["small_max_std"]) <= 5e-6
    with pytest.raises(RuntimeError):
        get_fom_max_std(ae, np.repeat(g["small_Zis"][:, :1], 3, axis=1))    # zero spread everywhere


def test_get_new_sample_point_end_to_end(tmp_path):
    """The whole greedy step through the reference-facing trainer API, against the oracle fed with
    the same GP draws (np.random seeded identically)."""
    phys = Physics()
    ps = ParamSpace()
    ae = _autoencoder(phys)
    ld = SINDy(5, phys.nt, {"sindy": {"fd_type": "sbp12"}})
    cfg = {"n_samples": 20, "lr": 1e-3, "n_iter": 1, "max_iter": 1, "max_greedy_iter": 1, "ld_weight": 0.1,
           "coef_weight": 1e-6, "path_checkpoint": str(tmp_path / "ckpt"), "path_results": str(tmp_path / "res")}
    trainer = BayesianGLaSDI(phys, ae, ld, ps, cfg)
    # best_coefs: calibrate encoded training trajectories (smooth synthetic series stand in for the FOM)
    Ztrain = torch.from_numpy(synth.make_latent_series(4, phys.nt, 5, phys.dt, seed=3))
    trainer.best_coefs = ld.calibrate(Ztrain, phys.dt, compute_loss=False, numpy=True)
    np.random.seed(0)
    new = trainer.get_new_sample_point()
    assert new.shape == (1, 2)

    # oracle with identical draws
    from gplasdi_b200.gp import sample_coefs_all
    Z0 = np.stack(initial_condition_latent(ps.test_space, phys, ae))
    gps = fit_gps(ps.train_space, trainer.best_coefs)
    np.random.seed(0)
    draws = sample_coefs_all(gps, ps.test_space, 20)
    Ws = [fc.weight.detach().cpu() for fc in ae.decoder.fcs]; bs = [fc.bias.detach().cpu() for fc in ae.decoder.fcs]
    idx, per, _ = orc.greedy_step(Ws, bs, "sigmoid", draws, Z0, phys.t_grid)
    assert np.array_equal(new, ps.test_space[idx].reshape(1, -1))
    assert np.max(np.abs(trainer.last_max_std - per) - 1e-5 * per) <= 5e-6


def test_fom_mean_std_matches_plot_prediction_statistics(golden_dir):
    """`fom_mean_std` = the `pred.mean(0)`, `pred.std(0)` of reference postprocess.plot_prediction (:32-34)."""
    g = np.load(os.path.join(golden_dir, "greedy_small.npz"))
    wl = synth.WORKLOADS["small"]
    phys = Physics(nt=wl.T, nx=wl.D)
    ae = _autoencoder(phys)
    Zis = g["small_Zis"][:2]
    mean, std = fom_mean_std(ae, Zis)
    assert mean.shape == std.shape == (2, wl.T, wl.D)
    Ws = [fc.weight.detach() for fc in ae.decoder.fcs]; bs = [fc.bias.detach() for fc in ae.decoder.fcs]
    X = orc.decoder_forward(Ws, bs, "sigmoid", torch.from_numpy(Zis[1].astype(np.float32))).numpy().astype(np.float64)
    assert np.max(np.abs(mean[1] - X.mean(0))) <= 5e-6 * np.max(np.abs(X))
    assert np.max(np.abs(std[1] - X.std(0))) <= 1e-5 * np.max(X.std(0)) + 5e-6


def test_sample_roms_and_average_rom_shapes():
    phys = Physics(nt=33, nx=64)
    ps = ParamSpace()
    ae = _autoencoder(phys, hidden=(16,))
    ld = SINDy(5, phys.nt, {"sindy": {"fd_type": "sbp12"}})
    rs = np.random.RandomState(1)
    gps = fit_gps(ps.train_space, rs.normal(size=(4, 30)))
    np.random.seed(3)
    Zis = sample_roms(ae, phys, ld, gps, ps.test_space[:5], 6)
    assert Zis.shape == (5, 6, 33, 5) and Zis.dtype == np.float64
    after = np.random.random_sample(3)
    # same seeded stream as the reference's loop: host draws (sklearn + np.random) give the same trajectories, and the
    # global generator is left where the reference would leave it
    from gplasdi_b200.gp import sample_coefs_all
    np.random.seed(3)
    draws = sample_coefs_all(gps, ps.test_space[:5], 6)
    np.testing.assert_array_equal(np.random.random_sample(3), after)
    z0h = np.stack(initial_condition_latent(ps.test_space[:5], phys, ae))
    Zh = ld.rollout_ensemble(draws, z0h, phys.t_grid).cpu().numpy()
    assert np.max(np.abs(Zis - Zh)) <= 1e-9 * np.max(np.abs(Zh))
    Zavg = average_rom(ae, phys, ld, gps, ps.test_space[:5])
    assert Zavg.shape == (5, 33, 5)
    mean, _ = eval_gp(gps, ps.test_space[:5])
    z0 = initial_condition_latent(ps.test_space[:5], phys, ae)
    ex = orc.simulate_exact(mean[2], z0[2], phys.dt, phys.nt)
    assert np.max(np.abs(Zavg[2] - ex)) <= 1e-9 * max(1.0, np.max(np.abs(ex)))


def test_c_host_program_on_the_c_abi(golden_dir, tmp_path):
    """tests/c_host/abi_smoke.c -- C99, the header and the CUDA runtime only -- runs the fused greedy step on the
    'small' fixture: same selected parameter and per-parameter values as the reference (through no Python at all)."""
    import struct
    import subprocess
    from test_boundary import _build_c_host
    from gplasdi_b200 import _lib
    exe = _build_c_host(tmp_path)
    g = np.load(os.path.join(golden_dir, "greedy_small.npz"))
    wl = synth.WORKLOADS["small"]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    inp = synth.make_inputs(wl)
    path = str(tmp_path / "problem.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<6i", len(Ws), _lib.ACT_IDS[wl.act], wl.P, wl.S, wl.n_z, wl.T))
        f.write(struct.pack("<d", wl.dt))
        f.write(np.asarray(wl.widths, dtype=np.int32).tobytes())
        for W, b in zip(Ws, bs):
            f.write(np.ascontiguousarray(W, dtype=np.float32).tobytes())
            f.write(np.ascontiguousarray(b, dtype=np.float32).tobytes())
        f.write(np.ascontiguousarray(inp["coefs"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(inp["z0"], dtype=np.float32).tobytes())
    out = subprocess.run([exe, path], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert int(lines[0].split()[1]) == int(g["small_m_index"][0])
    assert int(lines[0].split()[3]) > 0                                   # kernels were launched by the library
    per = np.array([float(x) for x in lines[1:]])
    ref = g["small_max_std"]
    assert per.shape == ref.shape and np.max(np.abs(per - ref) - 1e-5 * ref) <= 5e-6
