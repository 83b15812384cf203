#!/usr/bin/env python
"""Diagnostic: screened vs direct (3-pass / 1-pass) max_std per parameter for the trainer end-to-end scenario."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.gp import fit_gps, sample_coefs_all
from gplasdi_b200.latent_dynamics.sindy import SINDy
from gplasdi_b200.latent_space import initial_condition_latent
from gplasdi_b200.gplasdi import _engine_for
from test_gpu_api import Physics, ParamSpace, _autoencoder

phys = Physics(); ps = ParamSpace(); ae = _autoencoder(phys)
ld = SINDy(5, phys.nt, {"sindy": {"fd_type": "sbp12"}})
Ztrain = torch.from_numpy(synth.make_latent_series(4, phys.nt, 5, phys.dt, seed=3))
best = ld.calibrate(Ztrain, phys.dt, compute_loss=False, numpy=True)
Z0 = np.stack(initial_condition_latent(ps.test_space, phys, ae))
gps = fit_gps(ps.train_space, best)
np.random.seed(0)
draws = sample_coefs_all(gps, ps.test_space, 20)
eng = _engine_for(ae)
res = {}
for name, mode, passes in (("direct3", 0, 3), ("direct1", 0, 1), ("screened", 1, 3)):
    eng.set_option(_lib.OPT_MODE, mode); eng.set_option(_lib.OPT_SPLIT_PASSES, passes)
    ms, am, mv = eng.greedy_step(draws, Z0, phys.nt, phys.dt)
    res[name] = ms.cpu().numpy().astype(np.float64)
    try:
        eng.check_numeric()
    except Exception as e:
        print(name, "flag:", str(e)[:80])
    if mode == 1:
        print("screen stats", eng.screen_stats())
np.set_printoptions(linewidth=200, precision=4)
print("direct3 ", res["direct3"])
print("rel d1/d3-1", res["direct1"] / res["direct3"] - 1)
print("rel sc/d3-1", res["screened"] / res["direct3"] - 1)
print("coef draw std per param", draws.std(1).max(1))
