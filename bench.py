#!/usr/bin/env python
"""bench.py -- greedy-step ensemble rollouts/s (ODE + decode + variance + argmax) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1], the Burgers1D-shaped synthetic ensemble: 21x21 = 441 test
parameters x 1000 GP samples, nt = 501, nx = 1001, latent dim 5, decoder [5,100,1001] sigmoid.
A "step" is one pass of the whole greedy-sampling hot path over that ensemble:
rollout -> decode -> std over samples -> max over (t, x) -> argmax (-> max-loc all-reduce at N > 1).
Weak scaling (the headline line): every rank owns its own 441-parameter block of a (441 N)-point grid, no data-path
collective, one 8-byte max-loc all-reduce (NCCL) per step.

Printed JSON (rank 0, one line):
  value      rollouts/s with inputs resident in HBM, device-timed with CUDA events, max over ranks
  e2e        the same metric through the public Python API (Engine.checked = step + numeric check) from pinned HOST
             buffers, H2D of the draws and D2H of max_std + key inside the timed region
  roofline   the dominant kernel: tensor FLOP it ISSUES / its own device time / the measured dense tensor peak of
             MEASURED_PEAKS.json (burst figure for a sub-second timed region) -- a physical fraction; the algebraic gain of
             the covariance form over decoding every sample is a separate key, never a roofline fraction
  strong     ONE fixed grid (this workload's 441 parameters, the grid of the reference-run golden) split over the N ranks
             through sharding.sharded_greedy_step: ms/step, the selected global index and whether it equals the golden's
  configs    the other named shapes of BASELINE.json (configs[2..4]) at their named decoder / S / T on a fixed global block
             of test parameters split over the ranks, >= 1 s of timed work each at N = 1, with their own clock record
  cpu_baseline  the CPU oracle (port of the reference's own scipy/torch/numpy path) on a bounded sample on the host cores

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

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0: float, t1: float):
        """clocks / throttle reasons of the samples taken inside [t0, t1] (the sampler keeps running)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.lines):
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
# Blocks of test parameters the OTHER named shapes (BASELINE.json configs[2..4]) are timed on: a fixed GLOBAL block, split
# over the ranks like the real grid would be (strong scaling), sized for >= 1 s of timed device work at N = 1.
NAMED_BLOCKS = {"burgers2d": (None, 45), "vlasov1d1v": (64, 4), "rising_bubble": (48, 2)}     # (params or None = whole grid, steps)


def issued_tensor_flop(eng_widths, P, S, T, mode_gram: bool, passes: int):
    """Tensor-core FLOP the kernels ISSUE for P parameters (padding included): the physical numerator of a tensor-pipe
    fraction.  Covariance form (last-layer input width K <= 111): Gram MMAs (M = 128, N = ceil16(K + 1), K = 16 samples)
    + tf32 pre-activation MMAs (one front layer) in gram_kernel, the [P T x Q] x [Q x D] GEMM in quad_kernel.  Wide
    decoders: `passes` fp16 GEMM passes (1 = screening pass of the screened modes, 3 = fp32-faithful split of the direct
    mode) of the last front layer (xact_kernel) and of the last layer (xvar_kernel)."""
    L = eng_widths
    K, D = L[-2], L[-1]
    n_items = P * T
    if K + 1 <= 112 and mode_gram:
        n16 = -(-(K + 1) // 16) * 16
        gram = n_items * (-(-S // 16)) * 2.0 * 128 * n16 * 16
        if len(L) == 3:
            gram += n_items * (-(-S // 128)) * 3 * 2.0 * 128 * n16 * 8
        q_pad = -(-(K * (K + 1) // 2) // 64) * 64
        quad = -(-n_items // 128) * 128 * 2.0 * (-(-D // 256) * 256) * q_pad
        return gram, quad
    if K + 1 <= 112:
        return n_items * 2.0 * (-(-S // 128) * 128) * (-(-D // 128) * 128) * (-(-K // 16) * 16) * passes, 0.0
    Spad = -(-S // 256) * 256
    K1 = L[-3]
    xact = n_items * 2.0 * Spad * (-(-K // 128) * 128) * passes * (-(-K1 // 64) * 64)
    xvar = n_items * 2.0 * Spad * (-(-D // 128) * 128) * passes * (-(-K // 64) * 64)
    return xvar, xact


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
    # weak scaling: rank r owns block r of a (441 * world)-point grid -> its own seeded draws (rank 0 = the golden grid)
    from dataclasses import replace
    wl_r = replace(wl, seed=wl.seed + 17 * rank)
    inp = synth.make_inputs(wl_r)
    P, S, T, n = wl.P, wl.S, wl.T, wl.n_z
    index_offset = rank * P

    eng = Engine(local_rank)
    eng.set_decoder(Ws, bs, wl.act)
    screened = args.mode != "direct"
    mode_id = {"gram": _lib.MODE_SCREENED_GRAM, "screened": _lib.MODE_SCREENED, "direct": _lib.MODE_DIRECT}[args.mode]
    eng.set_option(_lib.OPT_MODE, mode_id)
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

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    total_ms = max_over_ranks(float(sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))))
    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    val_key, sel_idx = sharding.unpack_maxloc(int(key.item()))
    local_sel = int(out[1].item())                                        # this rank's own selection (rank 0: the golden grid)

    # ---- end to end through the public API from pinned host buffers -----------------------------
    # Engine.checked = greedy step + glasdi_check_numeric (+ the direct-mode repeat if the screening flag is up): the
    # call get_new_sample_point / sharded_greedy_step make, numeric check inside the timed region
    h2d = coefs_host.numel() * 8 + z0_host.numel() * 4
    d2h = P * 4 + 8 + 4 + 16
    host_ms = torch.empty(P, dtype=torch.float32).pin_memory()
    e2e_times = []
    barrier()
    for k in range(args.steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms, am, mv = eng.checked(lambda: eng.greedy_step(coefs_host, z0_host, T, wl.dt))   # H2D inside (pinned, overlapped)
        k2 = eng.pack_maxloc(mv, am, index_offset)
        sharding.allreduce_maxloc(k2)
        host_ms.copy_(ms, non_blocking=True)
        sel = int(k2.item())                                             # D2H read of the result (syncs)
        torch.cuda.synchronize()
        if k > 0:
            e2e_times.append(time.perf_counter() - t0)
    e2e_total = max_over_ranks(float(np.sum(e2e_times)))

    # ---- strong scaling of THIS grid: the 441 parameters of rank 0's (golden) grid split over the ranks ----------
    gold_path = os.path.join(ROOT, "tests", "golden", "greedy_named_burgers1d.npz")
    gold_idx = int(np.load(gold_path)["m_index"][0]) if os.path.exists(gold_path) else None
    bnd = sharding.shard_bounds(P, world)
    lo, hi = bnd[rank], bnd[rank + 1]
    inp0 = synth.make_inputs(wl, slice(lo, hi))                           # this rank's block of the UNSHIFTED (golden) grid
    c_s = torch.from_numpy(inp0["coefs"]).to(dev); z_s = torch.from_numpy(inp0["z0"]).to(dev)
    strong_ms = []
    skey = None
    for k in range(3 + args.steps):
        barrier()
        flush.fill_(k & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, skey = sharding.sharded_greedy_step(eng, c_s, z_s, T, wl.dt, lo)
        e1.record()
        e1.synchronize()
        if k >= 3:
            strong_ms.append(e0.elapsed_time(e1))
    strong_total = max_over_ranks(float(np.sum(strong_ms)))
    s_val, s_idx = sharding.unpack_maxloc(int(skey.item()))
    del c_s, z_s

    # ---- the other named shapes, each on its fixed global block split over the ranks -----------------------------
    named = []
    if not args.no_named:
        for nm, (blk, nsteps) in NAMED_BLOCKS.items():
            named.append(time_named_shape(eng, nm, blk, nsteps, world, rank, dev, sampler if rank == 0 else None,
                                          barrier, max_over_ranks, flush, mode_id))
        eng.set_decoder(Ws, bs, wl.act)
        eng.set_option(_lib.OPT_MODE, mode_id)
    if rank == 0:
        sampler.stop()

    if rank == 0:
        pk = peaks()
        rollouts_per_step = P * S * world
        value = rollouts_per_step * args.steps / (total_ms * 1e-3)
        flops_launch = wl.decoder_flops_per_rollout() * P * S          # algorithmic FLOP of one decode launch
        dec_avg_ms = float(np.mean(dec_ms)) / max(1, dec_launches / args.steps)
        step_ms = total_ms / args.steps
        # sub-second timed region -> the burst figure of MEASURED_PEAKS.json is the fair denominator
        peak = pk["burst"] if step_ms * args.steps < 1000.0 else pk["sustained"]
        peak_kind = "burst" if peak == pk["burst"] else "sustained"
        traffic, traffic_src = measured_traffic()
        issued, issued_2nd = issued_tensor_flop(wl.widths, P, S, T, args.mode == "gram", issued_passes)
        achieved = issued / (dec_avg_ms * 1e-3) / 1e12                 # ISSUED tensor FLOP of the dominant kernel / its time
        if args.mode == "gram":
            kernel = ("glasdi::g2::gram2_kernel<sigmoid, 5> (tcgen05 cta_group::1: tf32 pre-activation MMAs + f16 Gram MMAs on "
                      "K-major SW128 tiles, M=128 N=112 K=16; thread = hidden unit, Taylor-polynomial producers)")
            note = ("frac = tensor FLOP gram2_kernel ISSUES (padding included) / its own device time / peak: a physical "
                    "tensor-pipe fraction.  The kernel is bound by the CUDA-core work that feeds the tensor pipe -- P*S*T*K = "
                    "2.2e10 sigmoid evaluations (FMA-pipe polynomial around the pilot, exact MUFU form outside its radius) -- and by "
                    "the latency of the chain pre-activation MMA -> tcgen05.ld -> convert -> Gram MMA through one in-order tensor "
                    "pipe (profiles/r02_gram2_experiments.md), not by the pipe's throughput; quad_kernel (second GEMM) is tensor-bound. "
                    "The covariance form needs 5x fewer FLOP than decoding every sample (K^2 S + K(K+1)/2 D instead of K S D "
                    "per (parameter, time step)): that algebraic gain is reported separately as "
                    "algorithmic_speedup_vs_single_pass_roofline, never as a roofline fraction")
        else:
            kernel = "glasdi::dv2::decode_pair_kernel<VarT1, sigmoid, 5> (tcgen05 cta_group::2)"
            note = (f"frac = issued tensor FLOP ({issued_passes} fp16 pass(es), K 100->112, D 1001->1024, S 1000->1024 padding) / "
                    "kernel time / peak.  Decoding the ensemble into TMEM is bound by the accumulator drain (tcgen05.ld, "
                    "64 B/clk/SM): K/256 = 44 % of the tensor peak per pass")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (fp16 tensor-core screening pass over the ensemble, fp32 TMEM accumulate; exact fp32 decode + "
                      "fp64 sample sums at the near-maximum candidates; fp64 rollout)" if screened else
                      "f32 (fp16 split products x%d, fp32 TMEM accumulate; fp64 rollout)" % args.passes),
            "data": "synthetic",
            "config": {"workload": workload_desc(wl), "params_per_gpu": P, "global_params": P * world,
                       "mode": args.mode, "split_passes": issued_passes,
                       "parallelism": f"test-grid sharding x{world} + 1 max-loc all-reduce",
                       "l2": "256 MB buffer written between timed iterations; per-step working set 7.7 GB >> 126 MB L2",
                       "selected_global_index": sel_idx, "selected_index_rank0_grid": local_sel,
                       "golden_index": gold_idx,
                       "selected_index_matches_golden": (None if gold_idx is None else bool(local_sel == gold_idx))},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel,
                         "kernel_ms": dec_avg_ms, "kernel_share_of_step": dec_avg_ms / step_ms,
                         "rollout_kernel_ms": float(np.mean(roll_ms)),
                         "issued_flop_per_launch": issued,
                         "whole_step_issued_tflops": (issued + issued_2nd) / (step_ms * 1e-3) / 1e12,
                         "whole_step_issued_frac": (issued + issued_2nd) / (step_ms * 1e-3) / 1e12 / peak,
                         "algorithmic_flop_per_launch": flops_launch,
                         "algorithmic_speedup_vs_single_pass_roofline": flops_launch / (dec_avg_ms * 1e-3) / 1e12 / peak,
                         "peak_source": pk["source"] + f" bf16 dense, {peak_kind} figure ({'timed region < 1 s' if peak_kind == 'burst' else 'timed region >= 1 s'}); "
                                        "fp16 and bf16 share the tcgen05 kind::f16 rate",
                         "note": note},
            "e2e": {"value": rollouts_per_step * args.steps / e2e_total, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "call": "Engine.checked(greedy_step(pinned host draws)) + pack_maxloc + all-reduce + D2H of max_std and the key"},
            "strong": {"workload": workload_desc(wl) + " -- ONE fixed grid split over the ranks", "n_gpus": world,
                       "shard_bounds": bnd, "ms_per_step": strong_total / args.steps,
                       "rollouts_per_s": P * S * args.steps / (strong_total * 1e-3),
                       "selected_global_index": s_idx, "max_std": s_val, "golden_index": gold_idx,
                       "selected_index_matches_golden": (None if gold_idx is None else bool(s_idx == gold_idx)),
                       "call": "sharding.sharded_greedy_step (Engine.checked + pack_maxloc + NCCL max-loc all-reduce), device-timed, max over ranks"},
            "configs": named,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if args.mode == "gram":
            line["roofline"]["quad_kernel_ms"] = float(np.mean(quad_ms))
            line["roofline"]["quad_issued_flop_per_launch"] = issued_2nd
            line["roofline"]["quad_frac"] = issued_2nd / (float(np.mean(quad_ms)) * 1e-3) / 1e12 / peak
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
            ncpu = 4
            v, dt_cpu, threads = cpu_reference_rate(wl, Ws, bs, ncpu, seed_shift=216)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{ncpu} of 441 test parameters x {S} samples ({ncpu * S} rollouts, {dt_cpu:.1f} s): "
                                              f"odeint loop + time-chunked decode + numpy std; torch threads={threads}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from this round's `ncu --set full` capture of
    this same command (profiles/r02_gram2_kernel_traffic.json, written by tools/ncu_summary.py).  The capture records the
    sha1 of the kernel's source file: a capture of an older kernel reads as null, never as a stale number."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "r02_gram2_kernel_traffic.json")
    if not os.path.exists(path):
        return None, None
    try:
        d = json.load(open(path))
        src = os.path.join(ROOT, "gplasdi_b200", "csrc", "decode_gram2.cu")
        if d.get("source_sha1") != hashlib.sha1(open(src, "rb").read()).hexdigest():
            return None, "profiles/r02_gram2_kernel_traffic.json is of an older decode_gram2.cu: ignored"
        return float(d["dram_bytes_per_launch"]), "profiles/r02_gram2_kernel_traffic.json (" + d.get("how", "ncu --set full") + ")"
    except Exception:
        return None, None


def time_named_shape(eng, nm, blk, nsteps, world, rank, dev, sampler, barrier, max_over_ranks, flush, mode_id):
    """One named shape of BASELINE.json (configs[2..4]) on a fixed global block of its test grid, split over the ranks
    (strong scaling), through sharding.sharded_greedy_step; device-timed, max over ranks, clocks sampled in the window."""
    import torch
    from gplasdi_b200 import synth, _lib, sharding
    wl = synth.WORKLOADS[nm]
    Ws, bs = synth.make_decoder(wl.widths, wl.seed)
    eng.set_decoder(Ws, bs, wl.act)
    eng.set_option(_lib.OPT_MODE, mode_id)
    Pb = wl.P if blk is None else min(blk, wl.P)
    first = (wl.P - Pb) // 2
    bnd = sharding.shard_bounds(Pb, world)
    lo, hi = first + bnd[rank], first + bnd[rank + 1]
    inp = synth.make_inputs(wl, slice(lo, hi))
    coefs = torch.from_numpy(inp["coefs"]).to(dev); z0 = torch.from_numpy(inp["z0"]).to(dev)
    key = None
    for _ in range(2):
        _, key = sharding.sharded_greedy_step(eng, coefs, z0, wl.T, wl.dt, lo)
    barrier()
    l0 = eng.launch_count
    t0 = time.perf_counter()
    ms_list = []
    for k in range(nsteps):
        flush.fill_(k & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, key = sharding.sharded_greedy_step(eng, coefs, z0, wl.T, wl.dt, lo)
        e1.record()
        e1.synchronize()
        ms_list.append(e0.elapsed_time(e1))
    barrier()
    t1 = time.perf_counter()
    total = max_over_ranks(float(np.sum(ms_list)))
    val, idx = sharding.unpack_maxloc(int(key.item()))
    pk = peaks()
    peak = pk["burst"] if total < 1000.0 else pk["sustained"]
    sec = total * 1e-3 / nsteps
    alg = wl.decoder_flops_per_rollout() * Pb * wl.S
    wide_passes = 3 if mode_id == _lib.MODE_DIRECT else 1
    a, b = issued_tensor_flop(wl.widths, Pb, wl.S, wl.T, mode_id == _lib.MODE_SCREENED_GRAM, wide_passes)
    K = wl.widths[-2]
    path = ("gram_kernel (front MLP as a warp-level mma.sync chain with split fp16 operands + tcgen05 Gram MMAs) + quad_kernel + "
            "select/refine; bound by the producers (legacy HMMA + MUFU), not by the tcgen05 pipe: `frac` (issued tcgen05 FLOP) is "
            "small by construction, `frac_algorithmic` exceeds what a decode-every-sample kernel could reach because the "
            "covariance form needs K/D-times fewer FLOP" if K <= 111 else
            f"wide_pre_kernel (narrow front layers in registers) + xact_kernel / xvar_kernel ({wide_passes} fp16 pass(es) of the "
            "last front layer and of the last layer + variance epilogue on tcgen05, decode_wide.cu) + select + wide_cand_kernel "
            "(exact CUDA-core re-evaluation of the candidates' rows)")
    rec = {"workload": f"{nm}-shaped synthetic (BASELINE.json configs[{list(NAMED_BLOCKS).index(nm) + 2}]): P={wl.P} x S={wl.S}, "
                       f"T={wl.T}, D={wl.D}, decoder {wl.widths} {wl.act}",
           "params_timed": Pb, "first_param": first, "shard_bounds": bnd, "n_gpus": world, "steps": nsteps, "ms_per_step": sec * 1e3,
           "timed_ms_total": total, "rollouts_per_s": Pb * wl.S / sec, "seconds_per_full_grid_at_this_rate": sec * wl.P / Pb,
           "issued_tensor_tflops": (a + b) / sec / 1e12, "frac": (a + b) / sec / 1e12 / peak / world,
           "algorithmic_tflops": alg / sec / 1e12, "frac_algorithmic": alg / sec / 1e12 / peak / world, "peak": peak,
           "frac_definition": "frac = tensor FLOP the kernels issue (padding and split passes included) / whole-step device time / "
                              "(peak x n_gpus); frac_algorithmic = 2*sum(L_i L_i+1) decoder FLOP per rollout and time step / the same time / peak",
           "path": path, "selected_global_index": idx, "max_std": val, "gpu_launches": int(eng.launch_count - l0),
           "clocks": sampler.window(t0, t1) if sampler is not None else None}
    del coefs, z0
    return rec


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
    # one whole training iteration (reference gplasdi.py:177-199) at the shipped shapes: 25 training series of 1001 snapshots
    # x 1001 dofs through the [1001, 100, 5] autoencoder, reconstruction + SINDy + coefficient losses, backward -- every
    # kernel this package's own (forward/backward GEMMs, loss, fit, fit backward); optimiser step excluded
    try:
        from gplasdi_b200.latent_space import Autoencoder, mse_loss
        from gplasdi_b200.latent_dynamics.sindy import SINDy

        class _Phys:
            qgrid_size = [wl.D]
        torch.manual_seed(0)
        ae = Autoencoder(_Phys(), {"hidden_units": [100], "latent_dimension": n, "activation": wl.act}).cuda()
        ld = SINDy(n, 1001, {"sindy": {"fd_type": "sbp12", "coef_norm_order": "fro"}}, engine=eng)
        Xtr_t = torch.from_numpy(rs.normal(size=(ntr, 1001, wl.D)).astype(np.float32)).cuda()

        def train_iter():
            ae.zero_grad(set_to_none=True)
            Z = ae.encoder(Xtr_t)
            loss = mse_loss(ae.decoder(Z), Xtr_t, eng)
            _, l_ld, l_c = ld.calibrate(Z, 1e-3, compute_loss=True, numpy=True)
            (loss + 0.1 * l_ld / ntr + 1e-3 * l_c / ntr).backward()
        out["train_iteration_ms"] = timed(train_iter, 5)
        out["train_iteration_shape"] = [ntr, 1001, wl.D]
    except Exception as e:                                      # never let a context number break the contract line
        out["train_iteration_ms"] = None
        out["train_iteration_error"] = str(e)[:200]
    out["note"] = ("device times (CUDA events) of the SURVEY 8f kernels at this workload's shapes, outside the timed step: "
                   "GP posterior of 30 coefficient GPs at every test parameter, numpy-legacy MT19937/Box-Muller draws "
                   "reproduced on the GPU, batched encoder of the initial conditions, SINDy fit and its analytic backward, one "
                   "training iteration (autoencoder forward/backward + losses + calibration) on this package's kernels")
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
    ap.add_argument("--no-named", action="store_true", help="skip the per-config records of the other named shapes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)       # each step = 1 test parameter (~6 s of host work)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
