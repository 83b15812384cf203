"""MLP autoencoder with a CUDA (tcgen05) decoder; drop-in for the reference's latent-space module
(reference src/lasdi/latent_space.py:30-242).

Same constructor signatures, attribute names (`encoder`, `decoder`, `n_z`, `qgrid_size`) and
`state_dict` keys (`encoder.fcs.{i}.weight/bias`, `decoder.fcs.{i}.weight/bias`) as the reference,
so checkpoints written by either implementation load in the other.

`MultiLayerPerceptron.forward` runs on the hand-written CUDA kernels (glasdi_decode for the decoder,
glasdi_encode for the encoder) when it is used for inference -- no autograd graph requested -- which
is how the greedy-sampling path calls them (reference gplasdi.py:66, postprocess.py:32,
latent_space.py:57).  When gradients are required (the training loop) it evaluates with ordinary
differentiable torch ops, because the CUDA kernels have no backward.  `initial_condition_latent`
encodes all test parameters in ONE batched call instead of the reference's loop.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

# every parameter-free activation of the reference's act_dict (latent_space.py:5-28) has a CUDA implementation
# (include/glasdi_b200.h: enum glasdi_act); 'threshold' (two parameters) is torch-only: training works, inference raises
act_dict = {
    "sigmoid": torch.nn.Sigmoid, "tanh": torch.nn.Tanh, "softplus": torch.nn.Softplus, "ReLU": torch.nn.ReLU,
    "ELU": torch.nn.ELU, "SiLU": torch.nn.SiLU, "GELU": torch.nn.GELU,
    "hardshrink": torch.nn.Hardshrink, "hardsigmoid": torch.nn.Hardsigmoid, "hardtanh": torch.nn.Hardtanh,
    "hardswish": torch.nn.Hardswish, "leakyReLU": torch.nn.LeakyReLU, "logsigmoid": torch.nn.LogSigmoid,
    "ReLU6": torch.nn.ReLU6, "SELU": torch.nn.SELU, "CELU": torch.nn.CELU, "mish": torch.nn.Mish,
    "softshrink": torch.nn.Softshrink, "tanhshrink": torch.nn.Tanhshrink,
}


class _MlpTrain(torch.autograd.Function):
    """Training-mode MLP on the CUDA kernels: one autograd node instead of the per-layer graph torch records through
    MultiLayerPerceptron.forward (reference latent_space.py:157-164).  Saves the hidden layers' pre- and post-activations."""

    @staticmethod
    def forward(ctx, x, engine, act, n_linear, *params):
        weights = [p.detach().contiguous() for p in params[:n_linear]]
        biases = [p.detach().contiguous() for p in params[n_linear:]]
        xc = x.detach().contiguous()
        y, pre, post = engine.mlp_forward_train(xc, weights, biases, act)
        ctx.save_for_backward(xc, pre, post, *weights)
        ctx.engine, ctx.act, ctx.n_linear = engine, act, n_linear
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, pre, post, *weights = ctx.saved_tensors
        dx, dW, db = ctx.engine.mlp_backward(xc, weights, ctx.act, pre, post, gy.contiguous(), ctx.needs_input_grad[0])
        return (dx, None, None, None, *dW, *db)


class _MseLoss(torch.autograd.Function):
    """`MSELoss()(a, b)` with the loss and its gradient produced by ONE pass of glasdi_mse_loss_grad (the 'fused AE MSE' of
    SURVEY 8f rank 3); differentiable in its first argument (the reconstruction)."""

    @staticmethod
    def forward(ctx, a, b, engine):
        loss, grad = engine.mse_loss_grad(a.detach().contiguous(), b.detach().contiguous(), True)
        ctx.save_for_backward(grad)
        ctx.shape = a.shape
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).reshape(ctx.shape), None, None


def mse_loss(pred, target, engine=None):
    """Reconstruction loss of the training step (reference gplasdi.py:157,182): CUDA kernel when both live on the GPU."""
    if pred.is_cuda and target.is_cuda and pred.dtype == torch.float32 and target.dtype == torch.float32:
        from .engine import default_engine
        return _MseLoss.apply(pred, target, engine if engine is not None else default_engine(pred.device.index))
    return torch.nn.functional.mse_loss(pred, target)


def initial_condition_latent(param_grid, physics, autoencoder):
    """Z0[i] = encoder(u0(param_i)) as a list of fp32 numpy vectors (reference latent_space.py:30-63);
    all parameters are encoded in one batched pass instead of a python loop."""
    n_param = param_grid.shape[0]
    sol_shape = [1] + list(physics.qgrid_size)
    U0 = np.stack([np.asarray(physics.initial_condition(param_grid[i])).reshape(sol_shape) for i in range(n_param)])
    with torch.no_grad():
        enc = autoencoder.encoder
        if isinstance(enc, MultiLayerPerceptron):
            Z0 = enc(torch.Tensor(U0))                       # CUDA encoder kernel (glasdi_encode), one batch
        else:                                                # foreign module (e.g. the reference's own encoder)
            dev = next(autoencoder.parameters()).device
            Z0 = enc(torch.Tensor(U0).to(dev))
    Z0 = Z0[:, 0, :].detach().cpu().numpy()
    return [Z0[i] for i in range(n_param)]


class MultiLayerPerceptron(torch.nn.Module):
    def __init__(self, layer_sizes, act_type="sigmoid", reshape_index=None, reshape_shape=None,
                 threshold=0.1, value=0.0, num_heads=1):
        super().__init__()
        self.n_layers = len(layer_sizes)
        self.layer_sizes = list(layer_sizes)
        self.fcs = torch.nn.ModuleList(
            [torch.nn.Linear(layer_sizes[k], layer_sizes[k + 1]) for k in range(self.n_layers - 1)])
        self.init_weight()
        assert (reshape_index is None) or (reshape_index in [0, -1])
        assert (reshape_shape is None) or (np.prod(reshape_shape) == layer_sizes[reshape_index])
        self.reshape_index = reshape_index
        self.reshape_shape = None if reshape_shape is None else list(reshape_shape)
        self.act_type = act_type
        if act_type == "threshold":
            self.act = torch.nn.Threshold(threshold, value)
        elif act_type in act_dict:
            self.act = act_dict[act_type]()
        else:
            # PReLU / RReLU / multihead are excluded (learnable, random in train mode, broken upstream)
            raise ValueError(f"activation {act_type!r} is not supported")
        self._engine = None
        self._packed_version = None

    def init_weight(self):
        for fc in self.fcs:
            torch.nn.init.xavier_uniform_(fc.weight)

    # ---- CUDA inference path -------------------------------------------------------------------
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _require_cuda_decoder(self):
        """The inference path has no CPU fallback: fail loudly if the CUDA decoder cannot serve it."""
        if not torch.cuda.is_available():
            raise RuntimeError("gplasdi_b200 decoder inference needs a B200 GPU (no CPU fallback); "
                               "enable grad to evaluate with differentiable torch ops for training")
        if self.act_type not in _lib.ACT_IDS:
            raise NotImplementedError(f"activation {self.act_type!r} has no CUDA decoder kernel "
                                      f"(supported: {sorted(_lib.ACT_IDS)})")
        if self.n_layers < 3 or self.layer_sizes[0] > 8:
            raise NotImplementedError("CUDA decoder needs >= 1 hidden layer and latent dimension <= 8")

    def _require_cuda_encoder(self):
        if not torch.cuda.is_available():
            raise RuntimeError("gplasdi_b200 encoder inference needs a B200 GPU (no CPU fallback); "
                               "enable grad to evaluate with differentiable torch ops for training")
        if self.act_type not in _lib.ACT_IDS:
            raise NotImplementedError(f"activation {self.act_type!r} has no CUDA encoder kernel "
                                      f"(supported: {sorted(_lib.ACT_IDS)})")

    def _engine_with_encoder_weights(self):
        from .engine import default_engine
        if self._engine is None:
            self._engine = default_engine()
        ver = self._weights_version()
        if self._packed_version != ver or self._engine.encoder_owner is not self:
            self._engine.set_encoder([fc.weight for fc in self.fcs], [fc.bias for fc in self.fcs], self.act_type)
            self._engine.encoder_owner = self
            self._packed_version = ver
        return self._engine

    def _engine_with_weights(self):
        from .engine import default_engine
        if self._engine is None:
            self._engine = default_engine()
        ver = self._weights_version()
        if self._packed_version != ver or self._engine.decoder_owner is not self:
            self._engine.set_decoder([fc.weight for fc in self.fcs], [fc.bias for fc in self.fcs], self.act_type)
            self._engine.decoder_owner = self
            self._packed_version = ver
        return self._engine

    def forward(self, x):
        if self.reshape_index == 0:
            assert list(x.shape[-len(self.reshape_shape):]) == self.reshape_shape
            x = x.view(list(x.shape[:-len(self.reshape_shape)]) + [self.layer_sizes[0]])

        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.reshape_index != 0 and not needs_grad:
            # decoder inference = the hot path: hand-written CUDA kernel through the C ABI, or raise
            self._require_cuda_decoder()
            eng = self._engine_with_weights()
            y = eng.decode(x.detach().to(torch.float32)).to(x.device)
        elif not needs_grad:
            # encoder inference (initial_condition_latent): CUDA kernel chain through the C ABI, or raise
            self._require_cuda_encoder()
            eng = self._engine_with_encoder_weights()
            y = eng.encode(x.detach().to(torch.float32)).to(x.device)
        elif (x.is_cuda and self.act_type in _lib.ACT_IDS and all(p.is_cuda and p.dtype == torch.float32 for p in self.parameters())
              and x.dtype == torch.float32):
            # an autograd graph is being recorded on the GPU (the training step, reference gplasdi.py:180-181,197): this
            # package's forward / backward kernels behind one autograd node (glasdi_mlp_forward_train / glasdi_mlp_backward)
            from .engine import default_engine
            eng = self._engine if self._engine is not None else default_engine(x.device.index)
            lead = x.shape[:-1]
            y = _MlpTrain.apply(x.reshape(-1, self.layer_sizes[0]), eng, self.act_type, len(self.fcs),
                                *[fc.weight for fc in self.fcs], *[fc.bias for fc in self.fcs])
            y = y.reshape(*lead, self.layer_sizes[-1])
        else:
            # an autograd graph is being recorded on the host (CPU training / tests without a GPU): differentiable torch ops
            y = x
            for i in range(self.n_layers - 2):
                y = self.act(self.fcs[i](y))
            y = self.fcs[-1](y)

        if self.reshape_index == -1:
            y = y.view(list(y.shape[:-1]) + self.reshape_shape)
        return y


class Autoencoder(torch.nn.Module):
    """`Autoencoder(physics, config)` with config keys hidden_units, latent_dimension, activation
    (reference latent_space.py:191-242)."""

    def __init__(self, physics, config):
        super().__init__()
        self.qgrid_size = list(physics.qgrid_size)
        self.space_dim = int(np.prod(self.qgrid_size))
        hidden_units = list(config["hidden_units"])
        self.n_z = int(config["latent_dimension"])
        layer_sizes = [self.space_dim] + hidden_units + [self.n_z]
        act_type = config.get("activation", "sigmoid")
        threshold = config.get("threshold", 0.1)
        value = config.get("value", 0.0)
        self.encoder = MultiLayerPerceptron(layer_sizes, act_type, reshape_index=0, reshape_shape=self.qgrid_size,
                                            threshold=threshold, value=value)
        self.decoder = MultiLayerPerceptron(layer_sizes[::-1], act_type, reshape_index=-1,
                                            reshape_shape=self.qgrid_size, threshold=threshold, value=value)

    def forward(self, x):
        return self.decoder(self.encoder(x))

    def export(self):
        return {"autoencoder_param": self.cpu().state_dict()}

    def load(self, dict_):
        self.load_state_dict(dict_["autoencoder_param"])
