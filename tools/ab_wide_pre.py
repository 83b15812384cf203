#!/usr/bin/env python
"""A/B of the wide decoders' screening pass in ONE process, interleaved repetitions:
  A = GLASDI_OPT_WIDE_TENSOR 3: narrow front layers per thread on the CUDA cores (wide_pre_kernel),
  B = GLASDI_OPT_WIDE_TENSOR 2: the same chain on mma.sync with split fp16 operands (wide_pre_mma_kernel; default),
  D = GLASDI_MODE_DIRECT (three split passes, exact front layers): the fp32-faithful result both are compared with."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine

eng = Engine(0)
for nm, Pb in (("vlasov1d1v", 8), ("rising_bubble", 4), ("wide_k1000", 6)):
    wl = synth.WORKLOADS[nm]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    eng.set_decoder(Ws, bs, wl.act)
    Pb = min(Pb, wl.P)
    lo = (wl.P - Pb) // 2
    inp = synth.make_inputs(wl, slice(lo, lo + Pb))
    coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
    eng.set_option(_lib.OPT_MODE, _lib.MODE_DIRECT)
    ref = eng.greedy_step(coefs, z0, wl.T, wl.dt)[0].cpu().numpy()
    eng.check_numeric()
    eng.set_option(_lib.OPT_MODE, _lib.MODE_SCREENED_GRAM)
    times = {2: [], 3: []}
    res, stats = {}, {}
    for rep in range(4):
        for opt in (3, 2):
            eng.set_option(_lib.OPT_WIDE_TENSOR, opt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt); e1.record(); torch.cuda.synchronize()
            if rep: times[opt].append(e0.elapsed_time(e1))
            res[opt] = ms.cpu().numpy()
            stats[opt] = eng.screen_stats()
            eng.check_numeric()
    eng.set_option(_lib.OPT_WIDE_TENSOR, 2)
    print(nm, "P=%d" % Pb, "CUDA-core pre %.2f ms, mma.sync pre %.2f ms" % (np.median(times[3]), np.median(times[2])),
          "| max rel diff vs direct: A %.2e B %.2e" % tuple(float(np.max(np.abs(res[o] - ref) / np.maximum(ref, 1e-30))) for o in (3, 2)),
          "| screen stats (refine ms, cand, max dev) A", stats[3], "B", stats[2], flush=True)
