#!/usr/bin/env python
"""A/B of the wide-decoder paths in ONE process, interleaved repetitions, on blocks of the two wide-decoder shapes:
1 = GLASDI_MODE_DIRECT (three fp16 split passes of xact_kernel / xvar_kernel, fp32-faithful),
2 = default mode (ONE screening pass + exact CUDA-core re-evaluation of the candidates' rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine

eng = Engine(0)
for nm, Pb in (("vlasov1d1v", 4), ("rising_bubble", 2)):
    wl = synth.WORKLOADS[nm]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    eng.set_decoder(Ws, bs, wl.act)
    lo = (wl.P - Pb) // 2
    inp = synth.make_inputs(wl, slice(lo, lo + Pb))
    coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
    times = {1: [], 2: []}
    res = {}
    for rep in range(4):
        for mode in (1, 2):
            eng.set_option(_lib.OPT_MODE, _lib.MODE_DIRECT if mode == 1 else _lib.MODE_SCREENED_GRAM)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt); e1.record(); torch.cuda.synchronize()
            if rep: times[mode].append(e0.elapsed_time(e1))
            res[mode] = ms.cpu().numpy()
            eng.check_numeric()
            if mode == 2 and rep == 3: print("   screen stats (refine ms, candidates, max dev):", eng.screen_stats())
    print(nm, "P=%d" % Pb, {m: round(float(np.median(t)), 2) for m, t in times.items()}, "ms; max rel diff 1 vs 2:",
          float(np.max(np.abs(res[1] - res[2]) / np.maximum(res[1], 1e-30))), flush=True)
