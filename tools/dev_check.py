import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine
eng = Engine(0)
for nm, sl in (("small_softplus", None), ("burgers2d", slice(200, 232))):
    wl = synth.WORKLOADS[nm]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    eng.set_decoder(Ws, bs, wl.act)
    inp = synth.make_inputs(wl, sl)
    eng.set_option(_lib.OPT_MODE, _lib.MODE_DIRECT)
    ref = eng.greedy_step(inp["coefs"], inp["z0"], wl.T, wl.dt)[0].cpu().numpy()
    eng.set_option(_lib.OPT_MODE, _lib.MODE_SCREENED_GRAM)
    ms = eng.greedy_step(inp["coefs"], inp["z0"], wl.T, wl.dt)[0].cpu().numpy()
    try:
        eng.check_numeric(); fl = ""
    except Exception as e:
        fl = " FLAGGED"
    print(nm, "screen stats", eng.screen_stats(), "max rel diff vs direct", float(np.max(np.abs(ms - ref) / np.maximum(ref, 1e-30))), fl)
