#!/usr/bin/env python
"""Greedy step on EVERY named shape of BASELINE.json (configs[1..4]) on one B200, on a bounded block of test parameters
(the step is linear in P: parameters are independent), device-timed with CUDA events after warm-up.

One JSON line per shape: rollouts/s, the algorithmic decoder FLOP rate (2 * sum L_i L_{i+1} per rollout and
time step) against the measured dense tensor peak, and which kernel path served the decoder.  bench.py (the driver's
contract) measures configs[1] in full; this tool is the evidence for the other shapes (profiles/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine

BLOCKS = {"burgers1d": 441, "burgers2d": 128, "vlasov1d1v": 4, "rising_bubble": 2}


def main():
    pk = bench.peaks()
    eng = Engine(0)
    eng.set_option(_lib.OPT_STAGE_TIMING, 1)
    for nm, Pb in BLOCKS.items():
        wl = synth.WORKLOADS[nm]
        Ws, bs = synth.make_decoder(wl.widths, wl.seed)
        eng.set_decoder(Ws, bs, wl.act)
        Pb = min(Pb, wl.P)
        lo = (wl.P - Pb) // 2
        inp = synth.make_inputs(wl, slice(lo, lo + Pb))
        coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
        reps = 3
        eng.greedy_step(coefs, z0, wl.T, wl.dt)              # warm-up
        eng.check_numeric()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        flop = wl.decoder_flops_per_rollout() * Pb * wl.S
        K = wl.widths[-2]
        path = ("gram_kernel fast path (one front layer) + quad_kernel + refine" if K <= 111 and len(wl.widths) == 3 else
                "gram_kernel generic producers (deeper front MLP on the CUDA cores) + quad_kernel + refine" if K <= 111 else
                "early front layers on the CUDA cores (simt_fallback.cu) + last front layer (xact_kernel) and last layer + "
                "variance (xvar_kernel) as three-pass fp16-split tcgen05 GEMMs (decode_wide.cu)")
        print(json.dumps({
            "workload": f"{nm}-shaped synthetic (BASELINE.json configs[{list(BLOCKS).index(nm) + 1}]): P={wl.P} x S={wl.S}, T={wl.T}, "
                        f"D={wl.D}, decoder {wl.widths} {wl.act}", "params_in_block": Pb, "rollouts_per_s": Pb * wl.S / t,
            "seconds_per_full_grid": t * wl.P / Pb, "block_seconds": t,
            "algorithmic_tflops": flop / t / 1e12, "tensor_peak_tflops": pk["sustained"],
            "frac_of_tensor_roofline": flop / t / 1e12 / pk["sustained"], "path": path,
            "selected_index_in_block": int(am)}), flush=True)


if __name__ == "__main__":
    main()
