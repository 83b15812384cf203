#!/usr/bin/env python
"""bench.py -- greedy-step ensemble rollouts/s (ODE + decode + variance + argmax) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1], the Burgers1D-shaped synthetic ensemble: 21x21 = 441 test
parameters x 1000 GP samples, nt = 501, nx = 1001, latent dim 5, decoder [5,100,1001] sigmoid.
A "step" is one pass of the whole greedy-sampling hot path over that ensemble:
rollout -> decode -> std over samples -> max over (t, x) -> argmax (-> max-loc all-reduce at N > 1).
Weak scaling: every rank owns its own 441-parameter block of a (441 N)-point grid, no data-path
collective, one 8-byte max-loc all-reduce (NCCL) per step.

Printed JSON (rank 0, one line): metric/value = rollouts/s with inputs resident in HBM, device-timed
with CUDA events, max over ranks; `e2e` = the same metric through the public Python API from pinned
HOST buffers (H2D of the coefficient draws and D2H of max_std + index inside the timed region);
`roofline` = the decode+variance kernel's algorithmic FLOP rate against the measured dense bf16
tensor peak of MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle (port of the reference's own
scipy/torch/numpy path) timed on a bounded sample on this box's host cores.

--impl reference times that CPU path alone (the reference is pure Python + torch/scipy/numpy CPU
calls and cannot travel to the GPU box; oracle/reference_path.py restates it call for call and is
pinned to the reference by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "burgers1d"
METRIC = "greedy-step ensemble rollouts/s (ODE+decode+var+argmax)"
UNIT = "rollouts/s"


def workload_desc(wl):
    return (f"BASELINE configs[1] Burgers1D-shaped synthetic: P={wl.P} ({wl.grid[0]}x{wl.grid[1]}) x S={wl.S}, "
            f"T={wl.T}, D={wl.D}, latent {wl.n_z}, decoder {wl.widths} {wl.act}")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(wl, Ws, bs, n_params: int, seed_shift: int = 0):
    """rollouts/s of the CPU oracle (port of reference gplasdi.py:262-268) on `n_params` parameters of
    the workload: odeint double loop + time-chunked decode + numpy std + max.  Chunking over t is
    exact (std is per (t, dof)) and keeps the decoded slab at 64 MB instead of 2 GB per parameter."""
    import torch
    from gplasdi_b200 import synth
    from oracle import reference_path as orc
    inp = synth.make_inputs(wl, slice(seed_shift, seed_shift + n_params))
    t0 = time.perf_counter()
    Zis = orc.rollout_ensemble(inp["coefs"], inp["z0"], inp["t_grid"])
    idx, per = orc.fom_max_std(Ws, bs, wl.act, Zis, t_chunk=16)
    dt = time.perf_counter() - t0
    return n_params * wl.S / dt, dt, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on host cores; each step is a bounded
    sample (ONE test parameter = 1000 rollouts of the workload), extrapolation is linear in P."""
    import torch
    from gplasdi_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all the host threads torch can use: torchrun exports OMP_NUM_THREADS=1, which would halve the reference's rate
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    wl = synth.WORKLOADS[WORKLOAD]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    for w in range(args.warmup):
        cpu_reference_rate(wl, Ws, bs, 1, seed_shift=220 + w)
    times = []
    for k in range(args.steps):
        _, dt, threads = cpu_reference_rate(wl, Ws, bs, 1, seed_shift=230 + k)
        times.append(dt)
    total = float(np.sum(times))
    value = args.steps * wl.S / total
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(wl), "sample": "1 of 441 test parameters (1000 rollouts) per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x 1 test parameter x {wl.S} samples of the workload; "
                                   f"scipy odeint is single-threaded, torch Linear uses {torch.get_num_threads()} threads, "
                                   f"numpy std single-threaded; rate is per-parameter, linear in P"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gplasdi_b200 import synth, _lib, sharding
    from gplasdi_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    wl = synth.WORKLOADS[WORKLOAD]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    # weak scaling: rank r owns block r of a (441 * world)-point grid -> its own seeded draws
    from dataclasses import replace
    wl_r = replace(wl, seed=wl.seed + 17 * rank)
    inp = synth.make_inputs(wl_r)
    P, S, T, n = wl.P, wl.S, wl.T, wl.n_z
    index_offset = rank * P

    eng = Engine(local_rank)
    eng.set_decoder(Ws, bs, wl.act)
    screened = args.mode != "direct"
    eng.set_option(_lib.OPT_MODE, {"gram": _lib.MODE_SCREENED_GRAM, "screened": _lib.MODE_SCREENED,
                                   "direct": _lib.MODE_DIRECT}[args.mode])
    eng.set_option(_lib.OPT_SPLIT_PASSES, args.passes)
    eng.set_option(_lib.OPT_STAGE_TIMING, 1)
    issued_passes = 1 if screened else args.passes

    coefs_host = torch.from_numpy(inp["coefs"]).pin_memory()
    z0_host = torch.from_numpy(inp["z0"]).pin_memory()
    coefs = coefs_host.to(dev); z0 = z0_host.to(dev)
    out = (torch.empty(P, dtype=torch.float32, device=dev), torch.empty(1, dtype=torch.int64, device=dev),
           torch.empty(1, dtype=torch.float32, device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step():
        ms, am, mv = eng.greedy_step(coefs, z0, T, wl.dt, out=out)
        key = eng.pack_maxloc(mv, am, index_offset)
        sharding.allreduce_maxloc(key)
        return key

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        key = step()
    barrier()
    eng.check_numeric()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    dec_ms, roll_ms, ref_ms, quad_ms, dec_launches = [], [], [], [], 0
    n_cand, max_dev = 0, 0.0
    barrier()
    launches0 = eng.launch_count
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # L2 flush between timed iterations (untimed)
        ev0[k].record()
        key = step()
        ev1[k].record()
        ev1[k].synchronize()
        r, d, nl = eng.stage_times()
        roll_ms.append(r); dec_ms.append(d); dec_launches += nl
        if screened:
            rms, n_cand, max_dev = eng.screen_stats()
            ref_ms.append(rms); quad_ms.append(eng.last_quad_ms)
    barrier()
    t_wall1 = time.perf_counter()
    launches = eng.launch_count - launches0
    total_ms = float(sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)))
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    val_key, sel_idx = sharding.unpack_maxloc(int(key.item()))

    # ---- end to end through the public API from pinned host buffers -----------------------------
    h2d = coefs_host.numel() * 8 + z0_host.numel() * 4
    d2h = P * 4 + 8 + 4
    host_ms = torch.empty(P, dtype=torch.float32).pin_memory()
    e2e_times = []
    barrier()
    for k in range(args.steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms, am, mv = eng.greedy_step(coefs_host, z0_host, T, wl.dt)      # H2D inside (non_blocking, pinned)
        k2 = eng.pack_maxloc(mv, am, index_offset)
        sharding.allreduce_maxloc(k2)
        host_ms.copy_(ms, non_blocking=True)
        sel = int(k2.item())                                             # D2H read of the result (syncs)
        torch.cuda.synchronize()
        if k > 0:
            e2e_times.append(time.perf_counter() - t0)
    te = torch.tensor([float(np.sum(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_total = float(te.item())

    if rank == 0:
        pk = peaks()
        rollouts_per_step = P * S * world
        value = rollouts_per_step * args.steps / (total_ms * 1e-3)
        flops_launch = wl.decoder_flops_per_rollout() * P * S          # algorithmic FLOP of one decode launch
        dec_avg_ms = float(np.mean(dec_ms)) / max(1, dec_launches / args.steps)
        achieved = flops_launch / (dec_avg_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "decode_kernel_traffic.json")
        if os.path.exists(tpath) and args.mode == "gram":
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        K = wl.widths[-2]
        pad = 1.12 * 1024.0 / 1001.0 * 1024.0 / 1000.0
        if args.mode == "gram":
            n_items = P * T
            nmma = -(-S // 16)
            q_pad = -(-(K * (K + 1) // 2) // 64) * 64
            issued = n_items * nmma * 2.0 * 128 * (-(-(K + 1) // 16) * 16) * 16          # gram_kernel: Gram MMAs
            issued += n_items * (-(-S // 128)) * 3 * 2.0 * 128 * (-(-(K + 1) // 16) * 16) * 8   # + tf32 pre-activation MMAs
            issued_quad = -(-n_items // 128) * 128 * 2.0 * (-(-wl.D // 256) * 256) * q_pad   # quad_kernel MMAs
            kernel = ("glasdi::gr::gram_kernel<sigmoid, 5, fast> (tcgen05 cta_group::1: tf32 pre-activation MMAs + f16 Gram MMAs on "
                      "MN-major SW128 tiles, M=128 N=112 K=16)")
            note = ("frac is on the ALGORITHMIC decoder FLOP of the reference formulation (decode every sample, then std). "
                    "The covariance form reaches the same variances with 5x fewer FLOP (K^2 S + K(K+1)/2 D instead of K S D per "
                    "(parameter, time step)), so frac > 1 is algebra, not a faster tensor pipe: gram_kernel issues "
                    f"{issued:.3e} tensor FLOP (+ {issued_quad:.3e} in quad_kernel) and is bound by the CUDA-core work that "
                    "feeds it -- P*S*T*K = 2.2e10 sigmoid evaluations (MUFU ex2 + rcp: XU pipe 50 %, issue slots 55 % busy in "
                    "profiles/r01_gram_quad_kernels_bench_full_ncu.json).  kernel_ms is gram_kernel alone; quad_kernel_ms "
                    "is the second GEMM (80 % tensor-pipe active)")
        else:
            issued = flops_launch * issued_passes * pad
            issued_quad = 0.0
            kernel = "glasdi::dv2::decode_pair_kernel<VarT1, sigmoid, 5> (tcgen05 cta_group::2)"
            note = ("frac is on ALGORITHMIC decoder FLOP (single pass); the kernel issues "
                    f"{issued_passes} fp16 pass(es) per product and pads K 100->112, D 1001->1024, S 1000->1024: "
                    "issued_frac is the tensor-pipe view of the same launch.  Decoding the ensemble into TMEM is bound by "
                    "the accumulator drain (tcgen05.ld, 64 B/clk/SM): K/256 = 44 % of the tensor peak per pass")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (fp16 tensor-core screening pass over the ensemble, fp32 TMEM accumulate; exact fp32 decode + "
                      "fp64 sample sums at the near-maximum candidates; fp64 rollout)" if screened else
                      "f32 (fp16 split products x%d, fp32 TMEM accumulate; fp64 rollout)" % args.passes),
            "data": "synthetic",
            "config": {"workload": workload_desc(wl), "params_per_gpu": P, "global_params": P * world,
                       "mode": args.mode, "split_passes": issued_passes,
                       "parallelism": f"test-grid sharding x{world} + 1 max-loc all-reduce",
                       "l2": "256 MB buffer written between timed iterations; per-step working set 7.7 GB >> 126 MB L2",
                       "selected_global_index": sel_idx},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["sustained"], "traffic": traffic, "kernel": kernel,
                         "issued_frac": issued / (dec_avg_ms * 1e-3) / 1e12 / pk["sustained"],
                         "kernel_ms": dec_avg_ms, "rollout_kernel_ms": float(np.mean(roll_ms)),
                         "algorithmic_flop_per_launch": flops_launch, "issued_flop_per_launch": issued,
                         "peak_source": pk["source"] + " bf16 dense, sustained figure (kernel timed inside a long step); "
                                        "fp16 and bf16 share the tcgen05 kind::f16 rate",
                         "note": note},
            "e2e": {"value": rollouts_per_step * args.steps / e2e_total, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if args.mode == "gram":
            line["roofline"]["quad_kernel_ms"] = float(np.mean(quad_ms))
            line["roofline"]["quad_issued_flop_per_launch"] = issued_quad
        if screened:
            line["roofline"]["refine"] = {
                "select_plus_refine_ms": float(np.mean(ref_ms)), "candidates": int(n_cand),
                "max_rel_dev_screen_vs_exact": float(max_dev), "margin": 2.0 ** -eng.get_option(_lib.OPT_SCREEN_MARGIN),
                "note": "candidates = (t, dof) within the margin of a parameter's screened maximum, re-evaluated "
                        "exactly on the CUDA cores; their maximum is the reported max_std"}
        if world == 1:
            try:
                line["aux"] = aux_timings(eng, wl)
            except Exception as e:                                   # context only: never lose the headline line
                line["aux"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            ncpu = 8
            v, dt_cpu, threads = cpu_reference_rate(wl, Ws, bs, ncpu, seed_shift=216)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{ncpu} of 441 test parameters x {S} samples ({ncpu * S} rollouts, {dt_cpu:.1f} s): "
                                              f"odeint loop + time-chunked decode + numpy std; torch threads={threads}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def aux_timings(eng, wl):
    """Device times of the kernels either side of the greedy step, on this workload's shapes:
    GP posterior + legacy-stream coefficient draws, batched encoder, SINDy calibration forward / backward.
    Reported next to the headline for context; none of it is inside the timed step."""
    import torch
    from gplasdi_b200 import synth

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    rs = np.random.RandomState(3)
    P, S, n = wl.P, wl.S, wl.n_z
    nc = n * (n + 1)
    # 25 training points (4 corners + 21 greedy additions of the shipped Burgers run), RBF kernel matrices
    ntr = 25
    Xtr = rs.uniform([0.7, 0.9], [0.9, 1.1], size=(ntr, 2))
    a = np.linspace(0.7, 0.9, wl.grid[0]); w = np.linspace(0.9, 1.1, wl.grid[1])
    A, W = np.meshgrid(a, w, indexing="ij")
    grid = np.stack([A.ravel(), W.ravel()], axis=1)
    amp = rs.uniform(0.5, 5.0, nc); ls = rs.uniform(0.1, 1.0, nc)
    Ls, alphas = [], []
    for k in range(nc):
        d2 = ((Xtr[:, None, :] / ls[k] - Xtr[None, :, :] / ls[k]) ** 2).sum(-1)
        Kk = amp[k] * np.exp(-0.5 * d2) + 1e-10 * np.eye(ntr)
        Lk = np.linalg.cholesky(Kk)
        Ls.append(Lk); alphas.append(np.linalg.solve(Kk, rs.normal(size=ntr)))
    dev_args = [torch.as_tensor(x).cuda() for x in (Xtr, np.stack(alphas), np.stack(Ls), amp, ls, grid)]
    out["gp_posterior_ms"] = timed(lambda: eng.gp_predict(*dev_args), 10)
    mean, std = eng.gp_predict(*dev_args)
    st = np.random.RandomState(0).get_state()
    out["coef_draws_mt19937_ms"] = timed(lambda: eng.normal_mt19937(mean, std, S, state=st), 3)
    out["coef_draws_shape"] = [P, S, nc]
    We = [rs.normal(size=(100, wl.D)).astype(np.float32) * 0.03, rs.normal(size=(n, 100)).astype(np.float32) * 0.1]
    be = [np.zeros(100, np.float32), np.zeros(n, np.float32)]
    eng.set_encoder(We, be, wl.act)
    U = torch.from_numpy(rs.normal(size=(P, wl.D)).astype(np.float32)).cuda()
    out["encoder_ms"] = timed(lambda: eng.encode(U), 20)
    Zt = torch.from_numpy(synth.make_noisy_series(ntr, 1001, n, 1e-3, seed=5)).cuda()
    out["sindy_fit_us"] = 1e3 * timed(lambda: eng.sindy_fit(Zt, 1e-3, "sbp12", "fro"), 20)
    coefs = eng.sindy_fit(Zt, 1e-3, "sbp12", "fro")[0]
    ones = torch.ones(ntr, device="cuda")
    out["sindy_fit_backward_us"] = 1e3 * timed(lambda: eng.sindy_fit_backward(Zt, coefs, ones, ones, 1e-3, "sbp12", "fro"), 20)
    out["calibration_shape"] = [ntr, 1001, n]
    out["note"] = ("device times (CUDA events) of the SURVEY 8f kernels at this workload's shapes, outside the timed step: "
                   "GP posterior of 30 coefficient GPs at every test parameter, numpy-legacy MT19937/Box-Muller draws "
                   "reproduced on the GPU, batched encoder of the initial conditions, SINDy fit and its analytic backward")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--mode", choices=["gram", "screened", "direct"], default="gram",
                    help="gram (default): covariance-form fp16 screening + exact candidate re-evaluation; screened: "
                         "decoded-ensemble fp16 screening + the same re-evaluation; direct: --passes split passes")
    ap.add_argument("--passes", type=int, default=3, help="direct mode: fp16 split-product passes of the last layer (1..3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)       # each step = 1 test parameter (~6 s of host work)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
