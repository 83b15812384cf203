#!/usr/bin/env python
"""Device GP posterior + legacy-stream coefficient draws at the bench shape (441 x 1000 x 30) vs the host path."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gplasdi_b200.engine import Engine
from gplasdi_b200.gp import fit_gps, eval_gp, eval_gp_device, sample_coefs_all

P1, S, n = 21, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 30
rs = np.random.RandomState(11)
X = np.array([[0.7, 0.9], [0.7, 1.1], [0.9, 0.9], [0.9, 1.1]])
Y = rs.normal(size=(4, n))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    gps = fit_gps(X, Y)
a = np.linspace(0.7, 0.9, P1); w = np.linspace(0.9, 1.1, P1)
A, W = np.meshgrid(a, w, indexing="ij")
grid = np.stack([A.ravel(), W.ravel()], axis=1)
eng = Engine(0)
for _ in range(2):
    mean, std = eval_gp_device(gps, grid, eng)
torch.cuda.synchronize()
t0 = time.perf_counter(); mean, std = eval_gp_device(gps, grid, eng); torch.cuda.synchronize(); t1 = time.perf_counter()
t2 = time.perf_counter(); mh, sh = eval_gp(gps, grid); t3 = time.perf_counter()
print(f"posterior moments [{grid.shape[0]} x {n}]: device {1e3*(t1-t0):.2f} ms (incl. host packing), sklearn {1e3*(t3-t2):.1f} ms")
np.random.seed(0)
for _ in range(2):
    out, _ = eng.normal_mt19937(mean, std, S)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
np.random.seed(0)
ev[0].record(); out, _ = eng.normal_mt19937(mean, std, S); ev[1].record(); torch.cuda.synchronize()
dev_ms = ev[0].elapsed_time(ev[1])
np.random.seed(0)
t0 = time.perf_counter(); ref = sample_coefs_all(gps, grid, S); t1 = time.perf_counter()
pin = torch.from_numpy(ref).pin_memory()
torch.cuda.synchronize(); t2 = time.perf_counter(); d = pin.cuda(non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
err = np.max(np.abs(out.cpu().numpy() - ref)) / np.max(np.abs(ref))
print(f"draws [{grid.shape[0]} x {S} x {n}] = {ref.nbytes/1e6:.0f} MB: device {dev_ms:.2f} ms; host numpy {1e3*(t1-t0):.0f} ms "
      f"(+ sklearn predict) + H2D {1e3*(t3-t2):.1f} ms; max rel diff {err:.1e}")
