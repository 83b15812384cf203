#!/usr/bin/env python
"""HBM-roofline measurement of the SINDy calibration kernels:

  batched synthetic [B=4096, T=1001, n=5] fp32 series (82 MB in):
    glasdi_fd_dzdt   : algorithmic bytes = 4 B read + 4 B written per element  (164 MB per launch)
    glasdi_sindy_fit : algorithmic bytes = 4 B read per element + 30 coefs + 3 scalars per series
  and launch-latency figures (us per call) at the training-loop batch sizes B = 4..32.

Prints one JSON line per kernel; used for profiles/ and DESIGN.md, not by the driver.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gplasdi_b200.engine import Engine


def time_ms(fn, iters, flush=None, inner=8):
    """median / best device time per call; `inner` back-to-back launches per event pair keep the GPU queue
    full so that host launch overhead (tens of us per ctypes call) is not billed to ~50 us kernels.
    The input (82 MB) + output (82 MB) exceed nothing but are re-streamed from HBM when `flush` (256 MB
    written before every pair) evicted them from the 126 MB L2 -- only the first of the `inner` calls then
    sees a cold L2, so inner=1 is also reported for the cold-cache figure."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return float(np.median(ts)), float(np.min(ts))


def main():
    peak = 6550.4
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    eng = Engine(0)
    B, T, n = 8192, 1001, 5          # 164 MB in + 164 MB out: well beyond the 126 MB L2
    g = torch.Generator(device="cuda").manual_seed(0)
    Z = torch.randn(B, T, n, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for fd in ["sbp12", "sbp24", "sbp36", "sbp48"]:
        med, best = time_ms(lambda: eng.fd_dzdt(Z, 1e-3, fd), 10, None)
        byt = 2 * Z.numel() * 4
        out.append({"kernel": "fd_dzdt_kernel", "fd": fd, "shape": [B, T, n], "ms": med, "ms_best": best,
                    "algorithmic_bytes": byt, "achieved_gbs": byt / med / 1e6, "peak_gbs": peak,
                    "frac": byt / med / 1e6 / peak})
        med, best = time_ms(lambda: eng.sindy_fit(Z, 1e-3, fd, 1), 10, None)
        byt = Z.numel() * 4 + B * (n * (n + 1) + 3) * 4
        out.append({"kernel": f"sindy_fit_kernel<{n}>", "fd": fd, "shape": [B, T, n], "ms": med, "ms_best": best,
                    "algorithmic_bytes": byt, "achieved_gbs": byt / med / 1e6, "peak_gbs": peak,
                    "frac": byt / med / 1e6 / peak})
    for b in [4, 8, 16, 32]:
        Zs = Z[:b].contiguous()
        med, best = time_ms(lambda: eng.sindy_fit(Zs, 1e-3, "sbp12", 1), 30, None, inner=1)
        out.append({"kernel": f"sindy_fit_kernel<{n}>", "fd": "sbp12", "shape": [b, T, n], "us_per_call": med * 1e3,
                    "us_best": best * 1e3, "note": "training-loop batch size: launch/latency bound"})
    # backward of the fit (glasdi_sindy_fit_backward): reads Z, coefs, 2 upstream gradients per series, writes Zbar
    Zb = Z[:4096].contiguous()
    coefs = eng.sindy_fit(Zb, 1e-3, "sbp12", 1)[0]
    ones = torch.ones(Zb.shape[0], device="cuda")
    med, best = time_ms(lambda: eng.sindy_fit_backward(Zb, coefs, ones, ones, 1e-3, "sbp12", 1), 10, None)
    byt = 2 * Zb.numel() * 4 + Zb.shape[0] * (n * (n + 1) + 2) * 4
    out.append({"kernel": f"sindy_fit_bwd_kernel<{n}>", "fd": "sbp12", "shape": list(Zb.shape), "ms": med, "ms_best": best,
                "algorithmic_bytes": byt, "achieved_gbs": byt / med / 1e6, "peak_gbs": peak, "frac": byt / med / 1e6 / peak})
    for b in [4, 8, 16, 32]:
        Zs = Z[:b].contiguous()
        cs = eng.sindy_fit(Zs, 1e-3, "sbp12", 1)[0]
        o1 = torch.ones(b, device="cuda")
        med, best = time_ms(lambda: eng.sindy_fit_backward(Zs, cs, o1, o1, 1e-3, "sbp12", 1), 30, None, inner=1)
        out.append({"kernel": f"sindy_fit_bwd_kernel<{n}>", "fd": "sbp12", "shape": [b, T, n], "us_per_call": med * 1e3,
                    "us_best": best * 1e3, "note": "training-loop batch size: launch/latency bound"})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
