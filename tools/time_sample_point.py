#!/usr/bin/env python
"""Wall time of `BayesianGLaSDI.get_new_sample_point()` -- the reference-facing call (src/lasdi/gplasdi.py:241-274) -- at
the bench shape (21 x 21 test grid, 1000 samples, 501 steps, 1001 dofs, 25 training points), on 1 rank or under torchrun:

    python tools/time_sample_point.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/time_sample_point.py

Prints one JSON line per call from rank 0: total wall time and the stages (`last_timings`): Z0 by the batched encoder, the
scikit-learn GP fit (host, out of scope), GP posterior + numpy-legacy draws on the device, sharded greedy step + all-reduce."""
import contextlib, io, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist

from gplasdi_b200 import synth
from gplasdi_b200.gplasdi import BayesianGLaSDI
from gplasdi_b200.latent_dynamics.sindy import SINDy
from gplasdi_b200.latent_space import Autoencoder

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl")
wl = synth.WORKLOADS["burgers1d"]


class Physics:
    qgrid_size = [wl.D]; nt = wl.T; dt = wl.dt; t_grid = np.arange(wl.T) * wl.dt
    x_grid = np.linspace(-3.0, 3.0, wl.D)

    def initial_condition(self, param):
        a, w = param
        return a * np.exp(-self.x_grid ** 2 / 2 / w / w)


class ParamSpace:
    def __init__(self):
        a = np.linspace(0.7, 0.9, wl.grid[0]); w = np.linspace(0.9, 1.1, wl.grid[1])
        A, W = np.meshgrid(a, w, indexing="ij")
        self.test_space = np.stack([A.ravel(), W.ravel()], axis=1)
        rs = np.random.RandomState(1)
        self.train_space = np.concatenate([np.array([[0.7, 0.9], [0.7, 1.1], [0.9, 0.9], [0.9, 1.1]]),
                                           self.test_space[rs.choice(wl.P, 21, replace=False)]])
    def n_test(self): return self.test_space.shape[0]
    def n_train(self): return self.train_space.shape[0]


phys, ps = Physics(), ParamSpace()
torch.manual_seed(0)
ae = Autoencoder(phys, {"hidden_units": [100], "latent_dimension": 5, "activation": "sigmoid"}).cuda()
ld = SINDy(5, phys.nt, {"sindy": {"fd_type": "sbp12"}})
tmp = tempfile.mkdtemp()
cfg = {"n_samples": wl.S, "lr": 1e-3, "n_iter": 1, "max_iter": 1, "max_greedy_iter": 1, "ld_weight": 0.1, "coef_weight": 1e-6,
       "path_checkpoint": os.path.join(tmp, "c"), "path_results": os.path.join(tmp, "r"), "device": "cuda"}
tr = BayesianGLaSDI(phys, ae, ld, ps, cfg)
# best_coefs: the smooth mean-coefficient field of the synthetic workload at the training parameters (what a trained model's
# calibrations look like: the GP posterior is tight, as in a real run; random unrelated systems per training point would make
# the draws meaningless and trip the screening's deviation check into the direct mode)
grid01 = (ps.train_space - np.array([0.7, 0.9])) / 0.2
full = synth.make_inputs(synth.Workload("m", wl.grid, 1, 2, wl.widths, wl.act, wl.dt, wl.seed))
near = [int(np.argmin(np.sum((full["params"] - g) ** 2, axis=1))) for g in grid01]
tr.best_coefs = full["mean"][near]
for rep in range(3):
    np.random.seed(0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        new = tr.get_new_sample_point()
    wall = time.perf_counter() - t0
    fell_back = any("direct mode" in str(w.message) for w in wlist)
    if rank == 0:
        print(json.dumps({"call": "BayesianGLaSDI.get_new_sample_point", "n_gpus": world, "rep": rep, "wall_s": round(wall, 4),
                          "device_side_s": round(wall - tr.last_timings["fit_gps_host_s"], 4),
                          "stages": {k: round(v, 4) for k, v in tr.last_timings.items()},
                          "fell_back_to_direct_mode": fell_back, "selected": [float(x) for x in new.ravel()]}), flush=True)
if world > 1:
    dist.destroy_process_group()
