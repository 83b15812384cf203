#!/usr/bin/env python
"""One-shot GPU diagnostics (dev tool): runs every kernel on small seeded cases and prints the
gap to the oracle / golden fixtures.  Usage on the GPU box:  timeout 600 python tools/gpu_diag.py [stage ...]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from gplasdi_b200 import synth, _lib
from gplasdi_b200.engine import Engine
from oracle import reference_path as orc

G = os.path.join(ROOT, "tests", "golden")
stages = sys.argv[1:] or ["rollout", "calib", "greedy", "decode", "argmax", "big"]


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


eng = Engine(0)
print("engine ok; SMs:", torch.cuda.get_device_properties(0).multi_processor_count, flush=True)

if "rollout" in stages:
    g = np.load(os.path.join(G, "greedy_small.npz"))
    for nm in ["tiny", "small"]:
        wl = synth.WORKLOADS[nm]
        inp = synth.make_inputs(wl)
        for integ, sub in [(_lib.INT_EXPM, 1), (_lib.INT_RK4, 1), (_lib.INT_RK4, 4)]:
            eng.set_option(_lib.OPT_INTEGRATOR, integ); eng.set_option(_lib.OPT_RK4_SUBSTEPS, sub)
            Z = eng.rollout(inp["coefs"], inp["z0"], wl.T, wl.dt).cpu().numpy()
            print(f"rollout {nm} integ={integ} sub={sub}: rel vs reference odeint = {rel(Z, g[nm + '_Zis']):.3e}", flush=True)
        eng.set_option(_lib.OPT_INTEGRATOR, _lib.INT_EXPM); eng.set_option(_lib.OPT_RK4_SUBSTEPS, 1)
        ex = np.stack([[orc.simulate_exact(inp["coefs"][p, s], inp["z0"][p], wl.dt, wl.T) for s in range(wl.S)] for p in range(wl.P)])
        Z = eng.rollout(inp["coefs"], inp["z0"], wl.T, wl.dt).cpu().numpy()
        print(f"rollout {nm}: expm vs scipy expm = {rel(Z, ex):.3e}; odeint vs scipy expm = {rel(g[nm + '_Zis'], ex):.3e}", flush=True)

if "calib" in stages:
    g = np.load(os.path.join(G, "calibrate.npz"))
    for name in ["sbp12", "sbp24", "sbp36", "sbp48"]:
        for (B, T, order) in [(3, 65, 1), (2, 1001, "fro")]:
            dt = 1.0 / (T - 1)
            Z = synth.make_latent_series(B, T, 5, dt, seed=7)
            key = f"{name}_B{B}_T{T}"
            dz = eng.fd_dzdt(Z[0], dt, name).cpu().numpy()
            coefs, ls, lc, info = eng.sindy_fit(Z, dt, name, order)
            print(f"calib {key}: dzdt rel {rel(dz, g[key + '_dzdt0']):.2e} coefs rel {rel(coefs.cpu().numpy(), g[key + '_coefs']):.2e} "
                  f"loss_sindy {float(ls.sum()):.6e} vs {float(g[key + '_loss_sindy'][0]):.6e} "
                  f"loss_coef {float(lc.sum()):.6e} vs {float(g[key + '_loss_coef'][0]):.6e} info {info.cpu().numpy()}", flush=True)

if "greedy" in stages:
    g = np.load(os.path.join(G, "greedy_small.npz"))
    for nm in ["tiny", "small", "small_s200", "small_softplus"]:
        wl = synth.WORKLOADS[nm]
        Ws, bs = synth.make_decoder(wl.widths, wl.seed)
        inp = synth.make_inputs(wl)
        eng.set_decoder(Ws, bs, wl.act)
        for passes in [3, 2, 1]:
            eng.set_option(_lib.OPT_SPLIT_PASSES, passes)
            t0 = time.perf_counter()
            ms, am, mv = eng.greedy_step(inp["coefs"], inp["z0"], wl.T, wl.dt)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ms = ms.cpu().numpy()
            ref = g[nm + "_max_std"]
            relp = np.max(np.abs(ms - ref) / ref)
            print(f"greedy {nm} passes={passes}: idx {int(am)} (ref {int(g[nm + '_m_index'][0])}) "
                  f"max per-param rel err {relp:.3e}  [{(t1 - t0) * 1e3:.1f} ms]", flush=True)
            if passes == 3:
                print("   gpu:", ms[:6], "\n   ref:", ref[:6], flush=True)
        eng.set_option(_lib.OPT_SPLIT_PASSES, 3)
        ms2, am2, _ = eng.decode_max_std(g[nm + "_Zis"])
        print(f"decode_max_std {nm}: idx {int(am2)} rel {np.max(np.abs(ms2.cpu().numpy() - ref) / ref):.3e}", flush=True)
        eng.check_numeric()

if "decode" in stages:
    for nm in ["tiny", "small", "small_softplus"]:
        wl = synth.WORKLOADS[nm]
        Ws, bs = synth.make_decoder(wl.widths, wl.seed)
        eng.set_decoder(Ws, bs, wl.act)
        rs = np.random.RandomState(3)
        Z = rs.uniform(-1.5, 1.5, size=(7, 37, wl.n_z)).astype(np.float32)
        X = eng.decode(Z).cpu().numpy()
        Xr = orc.decoder_forward([torch.from_numpy(w) for w in Ws], [torch.from_numpy(b) for b in bs], wl.act,
                                 torch.from_numpy(Z)).numpy()
        print(f"decode {nm}: rel {rel(X, Xr):.3e}", flush=True)

if "argmax" in stages:
    v = torch.tensor([0.0, 0.5, 0.7, 0.7, 0.2])
    idx, val = eng.argmax_first(v)
    print("argmax_first", int(idx), float(val), "(expect 2, 0.7)")
    idx, val = eng.argmax_first(torch.zeros(5))
    print("argmax_first zeros", int(idx), "(expect -1)")
    key = eng.pack_maxloc(torch.tensor([0.7], device="cuda"), torch.tensor([2], device="cuda"), 100)
    print("pack", hex(int(key)))

if "big" in stages:
    wl = synth.WORKLOADS["burgers1d"]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    eng.set_decoder(Ws, bs, wl.act)
    P = 16
    inp = synth.make_inputs(wl, slice(0, P))
    coefs = torch.from_numpy(inp["coefs"]).cuda(); z0 = torch.from_numpy(inp["z0"]).cuda()
    for passes in [3, 2, 1]:
        eng.set_option(_lib.OPT_SPLIT_PASSES, passes)
        eng.greedy_step(coefs, z0, wl.T, wl.dt); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms, am, mv = eng.greedy_step(coefs, z0, wl.T, wl.dt)
        torch.cuda.synchronize()
        dt_ = time.perf_counter() - t0
        fl = wl.decoder_flops_per_rollout() * P * wl.S
        print(f"big P={P} passes={passes}: {dt_ * 1e3:.1f} ms  {P * wl.S / dt_:.3e} rollouts/s  alg {fl / dt_ / 1e12:.1f} TFLOP/s "
              f"idx {int(am)} ms[:4]={ms[:4].cpu().numpy()}", flush=True)
    eng.check_numeric()
print("diag done")
