#!/usr/bin/env python
"""A/B of the covariance-GEMM kernels on the bench workload (one process, interleaved): round-1 gram_kernel, gram2_kernel
with the Taylor-polynomial producers, gram2_kernel with the exact (MUFU) form everywhere.  Prints per-stage device times
(CUDA events inside the library), the screening deviation and the selected index of each."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine

nm = sys.argv[1] if len(sys.argv) > 1 else "burgers1d"
P = int(sys.argv[2]) if len(sys.argv) > 2 else None
wl = synth.WORKLOADS[nm]
Ws, bs = synth.make_decoder(wl.widths, wl.seed)
inp = synth.make_inputs(wl, None if P is None else slice(200, 200 + P))
eng = Engine(0)
eng.set_decoder(Ws, bs, wl.act)
eng.set_option(_lib.OPT_STAGE_TIMING, 1)
coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
variants = [("round1 gram_kernel", 0, 256), ("gram2 polynomial (radius 1.0)", 1, 256), ("gram2 exact form", 1, 0)]
acc = {v[0]: [] for v in variants}
ref = None
for rep in range(4):
    for name, kern, radius in variants:
        eng.set_option(_lib.OPT_GRAM_KERNEL, kern); eng.set_option(_lib.OPT_TAYLOR_RADIUS, radius)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt); e1.record(); torch.cuda.synchronize()
        eng.check_numeric()
        r, d, n = eng.stage_times(); rms, ncand, dev = eng.screen_stats()
        if ref is None:
            ref = ms.clone()
        if rep > 0:
            acc[name].append((e0.elapsed_time(e1), r, d, eng.last_quad_ms, rms))
        if rep == 3:
            a = np.mean(np.array(acc[name]), axis=0)
            print(json.dumps({"variant": name, "step_ms": a[0], "rollout_ms": a[1], "gram_ms": a[2], "quad_ms": a[3],
                              "select_refine_ms": a[4], "candidates": int(ncand), "max_screen_dev": float(dev), "selected": int(am),
                              "bit_equal_to_round1": bool(torch.equal(ms, ref))}), flush=True)
