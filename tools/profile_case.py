#!/usr/bin/env python
"""Short single-GPU run of the bench workload on a block of parameters, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 2          # 2 = covariance-form screening (default), 1 = x-space screening, 0 = direct
wl = synth.WORKLOADS[os.environ.get("GLASDI_WORKLOAD", "burgers1d")]
Ws, bs = synth.make_decoder(wl.widths, wl.seed)
inp = synth.make_inputs(wl, slice(200, 200 + P))
eng = Engine(0)
eng.set_decoder(Ws, bs, wl.act)
eng.set_option(_lib.OPT_SPLIT_PASSES, passes)
eng.set_option(_lib.OPT_STAGE_TIMING, 1)
eng.set_option(_lib.OPT_MODE, mode)
coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
for _ in range(reps):
    ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt)
    torch.cuda.synchronize()
    print("stage ms (rollout, decode, launches):", eng.stage_times(), "idx", int(am),
          "screen (refine ms, candidates, max dev):", eng.screen_stats() if mode >= 1 else None, flush=True)
eng.check_numeric()
if mode == 2:
    print("quad ms:", getattr(eng, "last_quad_ms", None))
