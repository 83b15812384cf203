#!/usr/bin/env python
"""Bench-shape run with a SMOOTH last decoder layer (trained decoders of a PDE field are smooth in the dof index):
how many (t, dof) candidates the screened modes re-evaluate, and what that costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wl = synth.WORKLOADS["burgers1d"]
Ws, bs = synth.make_decoder(wl.widths, wl.seed)
rs = np.random.RandomState(5)
D, K = Ws[-1].shape
xg = np.linspace(0.0, 1.0, D)[:, None]
modes = np.arange(1, 13)[None, :]
basis = np.cos(np.pi * modes * xg) / modes                  # [D, 12] smooth modes, decaying
Ws[-1] = (basis @ rs.normal(size=(12, K))).astype(np.float32) * 0.05
inp = synth.make_inputs(wl, slice(200, 200 + P))
eng = Engine(0)
eng.set_decoder(Ws, bs, wl.act)
eng.set_option(_lib.OPT_STAGE_TIMING, 1)
coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
for margin in (6, 8):
    eng.set_option(_lib.OPT_SCREEN_MARGIN, margin)
    for _ in range(2):
        ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt)
        torch.cuda.synchronize()
    r, d, n = eng.stage_times()
    rms, ncand, dev = eng.screen_stats()
    try:
        eng.check_numeric(); flag = ""
    except Exception as e:
        flag = " FLAGGED"
    print(f"margin 2^-{margin}: P={P} gram {d:.2f} ms quad {eng.last_quad_ms:.2f} ms refine {rms:.2f} ms candidates {ncand} "
          f"({ncand / P:.1f} per parameter) max dev {dev:.2e} idx {int(am)}" + flag, flush=True)
eng.set_option(_lib.OPT_MODE, _lib.MODE_DIRECT)
ms3 = eng.greedy_step(coefs, z0, wl.T, wl.dt)[0]
eng.set_option(_lib.OPT_MODE, _lib.MODE_SCREENED_GRAM)
print("max rel diff vs direct:", float(((ms - ms3).abs() / ms3).max()))
eng.set_option(_lib.OPT_MODE, _lib.MODE_SCREENED)
ms1 = eng.greedy_step(coefs, z0, wl.T, wl.dt)[0]; print("x-space screening:", eng.screen_stats(), "rel diff vs direct", float(((ms1 - ms3).abs() / ms3).max()))
